# round 2, call 13: where does the in-step GEMM slowdown come from?  schedule knobs A/B (C2 MFP, hybrid by-field encoder)
mkdir -p gpurun_out
i=0
for cfg in "MAP_B200_SINGLE_STREAM=1" "MAP_B200_NCE_SORT=early" "MAP_B200_NCE_SORT=early MAP_B200_MAIN_PRIO=0" "MAP_B200_TAB_PRIO=-1 MAP_B200_MAIN_PRIO=-2" "MAP_B200_MAIN_PRIO=-1"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 --timeline gpurun_out/r2_13_timeline_$i.txt > gpurun_out/r2_13_bench_$i.json 2> gpurun_out/r2_13_bench_$i.err; echo "[$i] $cfg rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/r2_13_bench_$i.json').read().strip().splitlines()[-1]); print('   value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))" || tail -5 gpurun_out/r2_13_bench_$i.err
done
