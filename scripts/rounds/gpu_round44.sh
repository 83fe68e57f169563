mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 100 python bench.py --steps 150 --warmup 10 --no-cpu-baseline --profile-steps 1 > gpurun_out/b44_$name.json 2> gpurun_out/b44_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/b44_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"; }
run late_hi_1 X=1
run early_hi_1 MAP_B200_NCE_SORT=early
run late_lo_1 MAP_B200_TAB_PRIO=0
run early_lo_1 MAP_B200_NCE_SORT=early MAP_B200_TAB_PRIO=0
run late_hi_2 X=1
run early_hi_2 MAP_B200_NCE_SORT=early
run late_lo_2 MAP_B200_TAB_PRIO=0
run early_lo_2 MAP_B200_NCE_SORT=early MAP_B200_TAB_PRIO=0
