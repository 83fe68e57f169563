# round 2, call 27: two-accumulator (terms = 6) tiles at <= 128 columns (main + corrections in one TMEM buffer): tests, levels, bench A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "bf16s" > gpurun_out/r2_27_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -2 gpurun_out/r2_27_pytest_gemm.log
LEVELS=F1,F2,S1 timeout 300 python scripts/bench_gemm_bf16s.py > gpurun_out/r2_27_levels_narrow.txt 2>&1; cat gpurun_out/r2_27_levels_narrow.txt
MAP_B200_DUAL_WIDE=1 LEVELS=F1,F2,S1 timeout 300 python scripts/bench_gemm_bf16s.py > gpurun_out/r2_27_levels_wide.txt 2>&1; grep -v "tile #" gpurun_out/r2_27_levels_wide.txt
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_27_$name.json 2> gpurun_out/r2_27_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_27_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), round(d['value']/1e6,3), 'e2e', round(d['e2e']['ms_per_step'],4), 'gemm us/launch', round(d['roofline']['us_per_launch'],1))" || tail -3 gpurun_out/r2_27_$name.err; }
run narrow X=1
run wide MAP_B200_DUAL_WIDE=1
timeout 900 python -m pytest tests/test_fullshape_gpu.py tests/test_model_gpu.py -m gpu -x -q > gpurun_out/r2_27_pytest_model.log 2>&1; echo "pytest model rc=$?"; tail -3 gpurun_out/r2_27_pytest_model.log
