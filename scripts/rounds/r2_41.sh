# round 2, call 41: sanity of HEAD after the scheduler-counter change: GEMM + model tests, short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_trainer_gpu.py -m gpu -x -q > gpurun_out/r2_41_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_41_pytest.log
timeout 200 python bench.py --steps 150 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_41_bench.json 2> gpurun_out/r2_41_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2_41_bench.json
