set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_dist_gpu.py -m gpu -q --timeout 900 > gpurun_out/t7_dist.log 2>&1; echo "dist rc=$?"; tail -n 15 gpurun_out/t7_dist.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/b7_mfp_2gpu.json 2> gpurun_out/b7_mfp_2gpu.err; echo "bench2 rc=$?"; head -c 900 gpurun_out/b7_mfp_2gpu.json; tail -n 12 gpurun_out/b7_mfp_2gpu.err
python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/b7_mfp_1gpu.json 2> gpurun_out/b7_mfp_1gpu.err; echo "bench1 rc=$?"; head -c 300 gpurun_out/b7_mfp_1gpu.json
