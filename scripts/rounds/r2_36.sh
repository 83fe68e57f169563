# round 2, call 36: sort grid size beside the dynamically scheduled GEMMs (the r2_20 experiment again: then every smaller grid lost)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 150 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_36_$name.json 2> gpurun_out/r2_36_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_36_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), round(d['value']/1e6,3), 'e2e', round(d['e2e']['ms_per_step'],4))" || tail -3 gpurun_out/r2_36_$name.err; }
run base X=1
run d148 MAP_B200_DEDUP_CTAS=148
run d96 MAP_B200_DEDUP_CTAS=96
run d64 MAP_B200_DEDUP_CTAS=64
run d32 MAP_B200_DEDUP_CTAS=32
run base2 X=1
