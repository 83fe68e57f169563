mkdir -p gpurun_out
run2() { # name, extra env...
  name=$1; shift
  env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 100 --warmup 10 --timeline gpurun_out/timeline38_${name}.txt > gpurun_out/b38_${name}.json 2> gpurun_out/b38_${name}.err; echo "bench $name rc=$?"; head -c 250 gpurun_out/b38_${name}.json; echo; grep -v "^\s*$" gpurun_out/b38_${name}.err | grep -iv "OMP_NUM\|\*\*\*\*" | tail -n 3
}
run2 mfp_2gpu_prio MAP_B200_NCCL_PRIO=1
run2 mfp_2gpu_noprio MAP_B200_NCCL_PRIO=0
