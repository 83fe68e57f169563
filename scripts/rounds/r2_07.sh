# round 2, call 7: whole GPU test suite at HEAD; SM partition experiment (GEMM clusters x sort CTAs) on the step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_07_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_07_pytest_gpu.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_07_$name.json 2> gpurun_out/r2_07_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_07_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), round(d['value']/1e6,3), 'e2e', round(d['e2e']['ms_per_step'],4))" || tail -3 gpurun_out/r2_07_$name.err; }
run base X=1
run c70_d16 MAP_B200_GEMM_CLUSTERS=70 MAP_B200_DEDUP_CTAS=16
run c66_d32 MAP_B200_GEMM_CLUSTERS=66 MAP_B200_DEDUP_CTAS=32
run c66_d64 MAP_B200_GEMM_CLUSTERS=66 MAP_B200_DEDUP_CTAS=64
run c74_d32 MAP_B200_DEDUP_CTAS=32
run c62_d48 MAP_B200_GEMM_CLUSTERS=62 MAP_B200_DEDUP_CTAS=48
run single_stream MAP_B200_SINGLE_STREAM=1
