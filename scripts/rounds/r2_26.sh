# round 2, call 26: TMA producer warp-uniform too, CTA-scope scheduler ops in the leader: tests, levels, bench; then ncu --set full with
# source of one forward level (F2) and one backward level (B2)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "bf16s" > gpurun_out/r2_26_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -2 gpurun_out/r2_26_pytest_gemm.log
LEVELS=F1,F2,B1,B2,B4,B5,S3,S4 timeout 300 python scripts/bench_gemm_bf16s.py > gpurun_out/r2_26_levels.txt 2>&1; cat gpurun_out/r2_26_levels.txt
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_26_$name.json 2> gpurun_out/r2_26_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_26_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), round(d['value']/1e6,3), 'e2e', round(d['e2e']['ms_per_step'],4), 'gemm us/launch', round(d['roofline']['us_per_launch'],1))" || tail -3 gpurun_out/r2_26_$name.err; }
run dynamic X=1
run static MAP_B200_GEMM_SCHED=static
LEVELS=F2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16s_kernel -s 1 -c 1 -o gpurun_out/r2_26_F2 -f python scripts/bench_gemm_bf16s.py > gpurun_out/r2_26_ncu_F2.log 2>&1; echo "ncu F2 rc=$?"; tail -3 gpurun_out/r2_26_ncu_F2.log
LEVELS=B2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16s_kernel -s 1 -c 1 -o gpurun_out/r2_26_B2 -f python scripts/bench_gemm_bf16s.py > gpurun_out/r2_26_ncu_B2.log 2>&1; echo "ncu B2 rc=$?"; tail -3 gpurun_out/r2_26_ncu_B2.log
ls -la gpurun_out/*.ncu-rep
