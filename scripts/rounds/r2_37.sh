# round 2, call 37: sort grid cap 96 as default, NCE forward at 3 CTAs per SM (80 registers), GEMM clusters 74 vs 72
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 150 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 3 > gpurun_out/r2_37_$name.json 2> gpurun_out/r2_37_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_37_$name.json').read().strip().splitlines()[-1]); k={x['kernel']:round(x['us_per_step'],1) for x in d['kernels']}; print('$name', round(d['ms_per_step'],4), round(d['value']/1e6,3), 'e2e', round(d['e2e']['ms_per_step'],4), 'nce', k.get('map_nce_fwd'), 'dedup', k.get('map_dedup_ids_ex'))" || tail -3 gpurun_out/r2_37_$name.err; }
run base X=1
run c72 MAP_B200_GEMM_CLUSTERS=72
run base2 X=1
timeout 200 python bench.py --task RFD --steps 150 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(\"rfd\", round(d[\"ms_per_step\"],4), round(d[\"value\"]/1e6,3))"
