mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 120 -k "gemm" > gpurun_out/t27_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -n 5 gpurun_out/t27_gemm.log
timeout 200 python scripts/bench_gemm_group.py > gpurun_out/gemm_group27_pair.txt 2> gpurun_out/gemm_group27.err; echo "group bench rc=$?"; cat gpurun_out/gemm_group27_pair.txt; tail -n 5 gpurun_out/gemm_group27.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/b27_mfp.json 2> gpurun_out/b27_mfp.err; echo "bench rc=$?"; head -c 230 gpurun_out/b27_mfp.json; tail -n 3 gpurun_out/b27_mfp.err
