# round 2, call 12: priority of the captured main branch (GEMM chain) vs side branches, x by-field encoder mode
mkdir -p gpurun_out
for cfg in "1 -1" "1 0" "0 -1" "2 -1" "1 -2"; do
  set -- $cfg; fe=$1; pr=$2
  MAP_B200_FIELD_ENC=$fe MAP_B200_MAIN_PRIO=$pr timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 --timeline gpurun_out/r2_12_timeline_fe${fe}_p${pr}.txt > gpurun_out/r2_12_bench_fe${fe}_p${pr}.json 2> gpurun_out/r2_12_bench_fe${fe}_p${pr}.err; echo "bench fe=$fe prio=$pr rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/r2_12_bench_fe${fe}_p${pr}.json').read().strip().splitlines()[-1]); print('fe=$fe prio=$pr value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['clocks'])" || tail -5 gpurun_out/r2_12_bench_fe${fe}_p${pr}.err
done
