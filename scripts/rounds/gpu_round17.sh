set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainer_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/t17_kernels.log 2>&1; echo "kernels+model rc=$?"; tail -n 4 gpurun_out/t17_kernels.log
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q --timeout 400 > gpurun_out/t17_dist.log 2>&1; echo "dist rc=$?"; tail -n 25 gpurun_out/t17_dist.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 50 --warmup 5 --timeline gpurun_out/timeline17_mfp_2gpu.txt > gpurun_out/b17_mfp_2gpu.json 2> gpurun_out/b17_mfp_2gpu.err; echo "bench2 rc=$?"; head -c 400 gpurun_out/b17_mfp_2gpu.json; grep -v "^\s*$" gpurun_out/b17_mfp_2gpu.err | grep -v "frame #" | tail -n 15
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 50 --warmup 5 --task RFD > gpurun_out/b17_rfd_2gpu.json 2> gpurun_out/b17_rfd_2gpu.err; echo "bench2 rfd rc=$?"; head -c 400 gpurun_out/b17_rfd_2gpu.json; tail -n 5 gpurun_out/b17_rfd_2gpu.err
