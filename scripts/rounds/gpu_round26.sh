mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/t26.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t26.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --timeline gpurun_out/timeline26_mfp.txt > gpurun_out/b26_mfp.json 2> gpurun_out/b26_mfp.err; echo "bench rc=$?"; head -c 230 gpurun_out/b26_mfp.json; tail -n 3 gpurun_out/b26_mfp.err
MAP_B200_GEMM_PAIR=0 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/b26_mfp_nopair.json 2> gpurun_out/b26_mfp_nopair.err; echo "bench nopair rc=$?"; head -c 230 gpurun_out/b26_mfp_nopair.json
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --task RFD > gpurun_out/b26_rfd.json 2> gpurun_out/b26_rfd.err; echo "bench rfd rc=$?"; head -c 230 gpurun_out/b26_rfd.json; tail -n 3 gpurun_out/b26_rfd.err
