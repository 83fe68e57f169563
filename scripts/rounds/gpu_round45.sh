mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/b45_default.json 2> gpurun_out/b45_default.err; echo "bench default rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/b45_default.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline'], d['clocks'])"; tail -n 2 gpurun_out/b45_default.err
