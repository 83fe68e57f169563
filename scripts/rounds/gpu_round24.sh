set -x
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 120 -k "gemm_group" > gpurun_out/t24_gemm.log 2>&1; echo "gemm group tests (pair) rc=$?"; tail -n 12 gpurun_out/t24_gemm.log
timeout 200 python scripts/bench_gemm_group.py > gpurun_out/gemm_group24_pair.txt 2> gpurun_out/gemm_group24.err; echo "group bench rc=$?"; cut -c1-200 gpurun_out/gemm_group24_pair.txt; tail -n 5 gpurun_out/gemm_group24.err
MAP_B200_GEMM_PAIR=0 timeout 200 python scripts/bench_gemm_group.py > gpurun_out/gemm_group24_nopair.txt 2> gpurun_out/gemm_group24n.err; echo "group bench nopair rc=$?"; cut -c1-200 gpurun_out/gemm_group24_nopair.txt
