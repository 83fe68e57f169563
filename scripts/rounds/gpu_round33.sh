mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/t33_dist.log 2>&1; echo "dist tests rc=$?"; tail -n 4 gpurun_out/t33_dist.log
run2() { # name, extra env...
  name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 100 --warmup 10 --timeline gpurun_out/timeline33_${name}.txt > gpurun_out/b33_${name}.json 2> gpurun_out/b33_${name}.err; echo "bench $name rc=$?"; head -c 250 gpurun_out/b33_${name}.json; echo; grep -v "^\s*$" gpurun_out/b33_${name}.err | grep -iv "OMP_NUM\|\*\*\*\*" | tail -n 6
}
run2 mfp_2gpu X=1
run2 mfp_2gpu_ncclbar MAP_B200_BARRIER=nccl
run2 mfp_2gpu_multi MAP_B200_DEDUP=multi MAP_B200_BENCH_VERBOSE=1
MAP_B200_DEDUP=multi MAP_B200_BENCH_VERBOSE=1 timeout 100 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/b33_mfp_multi_1gpu.json 2> gpurun_out/b33_mfp_multi_1gpu.err; echo "bench multi 1gpu rc=$?"; head -c 250 gpurun_out/b33_mfp_multi_1gpu.json; echo; tail -n 8 gpurun_out/b33_mfp_multi_1gpu.err
timeout 100 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/b33_mfp_1gpu.json 2> gpurun_out/b33_mfp_1gpu.err; echo "bench 1gpu rc=$?"; head -c 250 gpurun_out/b33_mfp_1gpu.json; echo
