# round 2, call 20: SM partition experiment repeated at HEAD (GEMM clusters x sort CTAs) on the C2 step
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_20_$name.json 2> gpurun_out/r2_20_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_20_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), round(d['value']/1e6,3), 'e2e', round(d['e2e']['ms_per_step'],4))" || tail -3 gpurun_out/r2_20_$name.err; }
run base X=1
run c74_d16 MAP_B200_DEDUP_CTAS=16
run c74_d32 MAP_B200_DEDUP_CTAS=32
run c74_d64 MAP_B200_DEDUP_CTAS=64
run c72_d16 MAP_B200_GEMM_CLUSTERS=72 MAP_B200_DEDUP_CTAS=16
run c70_d16 MAP_B200_GEMM_CLUSTERS=70 MAP_B200_DEDUP_CTAS=16
run c70_d32 MAP_B200_GEMM_CLUSTERS=70 MAP_B200_DEDUP_CTAS=32
run single_stream MAP_B200_SINGLE_STREAM=1
