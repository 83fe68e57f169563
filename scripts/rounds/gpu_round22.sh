set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 300 -k "gemm" > gpurun_out/t22_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -n 4 gpurun_out/t22_gemm.log
timeout 300 python scripts/bench_gemm_group.py > gpurun_out/gemm_group22.txt 2> gpurun_out/gemm_group22.err; echo "group bench rc=$?"; cut -c1-200 gpurun_out/gemm_group22.txt; tail -n 5 gpurun_out/gemm_group22.err
