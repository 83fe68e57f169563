# round 2, call 24: MMA issuer with warp-uniform control flow (no per-MMA R2UR waterfall): kernel tests, levels, step bench vs static and vs the previous kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "bf16s" > gpurun_out/r2_24_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -2 gpurun_out/r2_24_pytest_gemm.log
LEVELS=F1,F2,B1,B2,B4,B5,S1,S3,S4 timeout 300 python scripts/bench_gemm_bf16s.py > gpurun_out/r2_24_levels_dynamic.txt 2>&1; cat gpurun_out/r2_24_levels_dynamic.txt
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_24_$name.json 2> gpurun_out/r2_24_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_24_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), round(d['value']/1e6,3), 'e2e', round(d['e2e']['ms_per_step'],4), 'gemm us/launch', round(d['roofline']['us_per_launch'],1))" || tail -3 gpurun_out/r2_24_$name.err; }
run dynamic X=1
run static MAP_B200_GEMM_SCHED=static
cp map_code_b200/libmap_b200.so /tmp/new.so; cp build_old/libmap_b200.so map_code_b200/libmap_b200.so
run old X=1
cp /tmp/new.so map_code_b200/libmap_b200.so
run dynamic2 X=1
