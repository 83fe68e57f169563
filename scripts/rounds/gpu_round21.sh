set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 300 -k "gemm" > gpurun_out/t21_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -n 15 gpurun_out/t21_gemm.log
timeout 300 python scripts/bench_gemm_group.py > gpurun_out/gemm_group21.txt 2> gpurun_out/gemm_group21.err; echo "group bench rc=$?"; cat gpurun_out/gemm_group21.txt; tail -n 5 gpurun_out/gemm_group21.err
