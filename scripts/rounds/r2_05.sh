# round 2, call 5: BLOCK_N planned per launch from a makespan model; kernel tests, level timing, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "bf16s" > gpurun_out/r2_05_bf16s.log 2>&1; rc=$?; echo "bf16s rc=$rc"; tail -4 gpurun_out/r2_05_bf16s.log
if [ $rc -ne 0 ]; then exit 0; fi
LEVELS=F1,F2,F4,B1,B2,B4,B5 timeout 600 python scripts/bench_gemm_bf16s.py > gpurun_out/r2_05_gemm_levels.txt 2> gpurun_out/r2_05_gemm_levels.err; echo "levels rc=$?"; grep -v "tile #" gpurun_out/r2_05_gemm_levels.txt; tail -3 gpurun_out/r2_05_gemm_levels.err
timeout 600 python bench.py --steps 100 --warmup 10 --profile-steps 2 --no-secondary --no-cpu-baseline > gpurun_out/r2_05_bench_mfp.json 2> gpurun_out/r2_05_bench_mfp.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_05_bench_mfp.err
python -c "import json;d=json.loads(open('gpurun_out/r2_05_bench_mfp.json').read().strip().splitlines()[-1]);print(json.dumps({k:d[k] for k in ('value','ms_per_step')},indent=1)); print(d['roofline']['us_per_launch'], d['roofline']['tensor_pipe_tflops_issued'])"
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_trainer_gpu.py tests/test_fullshape_gpu.py -q > gpurun_out/r2_05_model.log 2>&1; echo "model rc=$?"; tail -6 gpurun_out/r2_05_model.log
