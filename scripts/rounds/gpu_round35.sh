mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m pytest tests/test_dist_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/t35_dist.log 2>&1; echo "dist tests rc=$?"; tail -n 4 gpurun_out/t35_dist.log
run2() { # name, task, extra env...
  name=$1; shift; task=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 100 --warmup 10 --task $task --timeline gpurun_out/timeline35_${name}.txt > gpurun_out/b35_${name}.json 2> gpurun_out/b35_${name}.err; echo "bench $name rc=$?"; head -c 250 gpurun_out/b35_${name}.json; echo; grep -v "^\s*$" gpurun_out/b35_${name}.err | grep -iv "OMP_NUM\|\*\*\*\*" | tail -n 6
}
run2 mfp_2gpu MFP X=1
run2 rfd_2gpu RFD X=1
MAP_B200_DEDUP=multi MAP_B200_BENCH_VERBOSE=1 timeout 70 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/b35_mfp_multi_1gpu.json 2> gpurun_out/b35_mfp_multi_1gpu.err; echo "bench multi 1gpu rc=$?"; head -c 250 gpurun_out/b35_mfp_multi_1gpu.json; echo; tail -n 40 gpurun_out/b35_mfp_multi_1gpu.err
