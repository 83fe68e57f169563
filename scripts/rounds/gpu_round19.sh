set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/t19_kernels.log 2>&1; echo "kernels+model rc=$?"; tail -n 3 gpurun_out/t19_kernels.log
timeout 300 python scripts/trace_gemm.py > gpurun_out/trace_gemm19.txt 2> gpurun_out/trace_gemm19.err; echo "trace rc=$?"; cut -c1-100 gpurun_out/trace_gemm19.txt; tail -n 3 gpurun_out/trace_gemm19.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --timeline gpurun_out/timeline19_mfp.txt > gpurun_out/b19_mfp.json 2> gpurun_out/b19_mfp.err; echo "bench rc=$?"; head -c 330 gpurun_out/b19_mfp.json; tail -n 3 gpurun_out/b19_mfp.err
MAP_B200_PDL=0 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/b19_mfp_nopdl.json 2> gpurun_out/b19_mfp_nopdl.err; echo "bench nopdl rc=$?"; head -c 330 gpurun_out/b19_mfp_nopdl.json
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --task RFD > gpurun_out/b19_rfd.json 2> gpurun_out/b19_rfd.err; echo "bench rfd rc=$?"; head -c 330 gpurun_out/b19_rfd.json; tail -n 3 gpurun_out/b19_rfd.err
