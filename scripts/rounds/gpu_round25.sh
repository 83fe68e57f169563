mkdir -p gpurun_out
timeout 200 python scripts/bench_gemm_group.py > gpurun_out/gemm_group25_pair.txt 2> gpurun_out/gemm_group25.err; echo "group bench rc=$?"; cat gpurun_out/gemm_group25_pair.txt; tail -n 5 gpurun_out/gemm_group25.err
MAP_B200_GEMM_PAIR=0 timeout 200 python scripts/bench_gemm_group.py > gpurun_out/gemm_group25_nopair.txt 2> gpurun_out/gemm_group25n.err; echo "group bench nopair rc=$?"; cat gpurun_out/gemm_group25_nopair.txt
