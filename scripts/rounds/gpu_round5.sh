set -x
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 900 > gpurun_out/t5_kernels.log 2>&1; echo "kernels rc=$?"; tail -n 4 gpurun_out/t5_kernels.log
python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 900 > gpurun_out/t5_model.log 2>&1; echo "model rc=$?"; tail -n 6 gpurun_out/t5_model.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --dump-profile gpurun_out/p5_shapes.txt > gpurun_out/b5_mfp.json 2> gpurun_out/b5_mfp.err; echo "bench rc=$?"; head -c 600 gpurun_out/b5_mfp.json; tail -n 3 gpurun_out/b5_mfp.err
python scripts/tune_gemm.py > gpurun_out/tune5.txt 2>&1; echo "tune rc=$?"; cat gpurun_out/tune5.txt
