mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_dist_gpu.py -m gpu -q -x --timeout 150 -k "MFP and p2p" > gpurun_out/t36_dist.log 2>&1; echo "dist tests rc=$?"; tail -n 4 gpurun_out/t36_dist.log
run2() { # name, task, extra env...
  name=$1; shift; task=$1; shift
  env "$@" MAP_B200_BENCH_VERBOSE=1 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 100 --warmup 10 --task $task --timeline gpurun_out/timeline36_${name}.txt > gpurun_out/b36_${name}.json 2> gpurun_out/b36_${name}.err; echo "bench $name rc=$?"; head -c 250 gpurun_out/b36_${name}.json; echo; grep -v "^\s*$" gpurun_out/b36_${name}.err | grep -iv "OMP_NUM\|\*\*\*\*" | tail -n 12
}
run2 mfp_2gpu MFP X=1
run2 rfd_2gpu RFD X=1
