set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/t12_kernels.log 2>&1; echo "kernels rc=$?"; tail -n 5 gpurun_out/t12_kernels.log
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_trainer_gpu.py -m gpu -q --timeout 300 > gpurun_out/t12_model.log 2>&1; echo "model rc=$?"; tail -n 5 gpurun_out/t12_model.log
timeout 300 python scripts/trace_gemm.py > gpurun_out/trace_gemm12.txt 2> gpurun_out/trace_gemm12.err; echo "trace rc=$?"; cut -c1-400 gpurun_out/trace_gemm12.txt; tail -n 3 gpurun_out/trace_gemm12.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --dump-profile gpurun_out/p12_shapes.txt --timeline gpurun_out/timeline12_mfp.txt > gpurun_out/b12_mfp.json 2> gpurun_out/b12_mfp.err; echo "bench rc=$?"; head -c 400 gpurun_out/b12_mfp.json; tail -n 3 gpurun_out/b12_mfp.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --task RFD --timeline gpurun_out/timeline12_rfd.txt > gpurun_out/b12_rfd.json 2> gpurun_out/b12_rfd.err; echo "bench rfd rc=$?"; head -c 400 gpurun_out/b12_rfd.json; tail -n 3 gpurun_out/b12_rfd.err
timeout 900 python scripts/tune_gemm.py > gpurun_out/tune12.txt 2> gpurun_out/tune12.err; echo "tune rc=$?"; tail -n 3 gpurun_out/tune12.err
