# round 2, call 22: dynamic tile scheduler of the split-bf16 GEMM: kernel tests, per-level timing dynamic vs static, step bench both ways
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "bf16s" > gpurun_out/r2_22_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -4 gpurun_out/r2_22_pytest_gemm.log
LEVELS=F1,F2,B1,B2,B4,B5 timeout 300 python scripts/bench_gemm_bf16s.py > gpurun_out/r2_22_levels_dynamic.txt 2>&1; grep -v "tile #" gpurun_out/r2_22_levels_dynamic.txt
MAP_B200_GEMM_SCHED=static LEVELS=F1,F2,B1,B2,B4,B5 timeout 300 python scripts/bench_gemm_bf16s.py > gpurun_out/r2_22_levels_static.txt 2>&1; grep -v "tile #" gpurun_out/r2_22_levels_static.txt
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_22_$name.json 2> gpurun_out/r2_22_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_22_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), round(d['value']/1e6,3), 'e2e', round(d['e2e']['ms_per_step'],4), 'gemm us/launch', round(d['roofline']['us_per_launch'],1))" || tail -3 gpurun_out/r2_22_$name.err; }
run dynamic X=1
run static MAP_B200_GEMM_SCHED=static
timeout 900 python -m pytest tests/test_fullshape_gpu.py tests/test_model_gpu.py -m gpu -x -q > gpurun_out/r2_22_pytest_model.log 2>&1; echo "pytest model rc=$?"; tail -4 gpurun_out/r2_22_pytest_model.log
