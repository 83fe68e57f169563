set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python scripts/trace_gemm.py > gpurun_out/trace_gemm10.txt 2> gpurun_out/trace_gemm10.err; echo "trace rc=$?"; cat gpurun_out/trace_gemm10.txt; tail -n 3 gpurun_out/trace_gemm10.err
timeout 400 python -m pytest tests/test_dist_gpu.py -m gpu -q --timeout 300 > gpurun_out/t10_dist.log 2>&1; echo "dist rc=$?"; tail -n 5 gpurun_out/t10_dist.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/b10_mfp_2gpu.json 2> gpurun_out/b10_mfp_2gpu.err; echo "bench2 rc=$?"; head -c 1500 gpurun_out/b10_mfp_2gpu.json; tail -n 12 gpurun_out/b10_mfp_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 50 --warmup 5 --task RFD > gpurun_out/b10_rfd_2gpu.json 2> gpurun_out/b10_rfd_2gpu.err; echo "bench2 rfd rc=$?"; head -c 600 gpurun_out/b10_rfd_2gpu.json; tail -n 5 gpurun_out/b10_rfd_2gpu.err
