# round 2, call 4: 8 epilogue warps + cost-ordered snake schedule: kernel tests, per-level timeline, step bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "bf16s" > gpurun_out/r2_04_bf16s.log 2>&1; rc=$?; echo "bf16s rc=$rc"; tail -4 gpurun_out/r2_04_bf16s.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python scripts/bench_gemm_bf16s.py > gpurun_out/r2_04_gemm_levels.txt 2> gpurun_out/r2_04_gemm_levels.err; echo "levels rc=$?"; cat gpurun_out/r2_04_gemm_levels.txt; tail -3 gpurun_out/r2_04_gemm_levels.err
timeout 600 python bench.py --steps 100 --warmup 10 --profile-steps 2 --no-secondary --no-cpu-baseline --timeline gpurun_out/r2_04_timeline.txt > gpurun_out/r2_04_bench_mfp.json 2> gpurun_out/r2_04_bench_mfp.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_04_bench_mfp.err
python -c "import json;d=json.loads(open('gpurun_out/r2_04_bench_mfp.json').read().strip().splitlines()[-1]);print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','roofline')},indent=1)); [print(k) for k in d['kernels']]"
cat gpurun_out/r2_04_timeline.txt
