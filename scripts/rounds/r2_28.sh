# round 2, call 28: early dense AdamW (layers >= 1 + head under the last two GEMM levels): model tests, bench A/B, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_trainer_gpu.py tests/test_fullshape_gpu.py -m gpu -x -q > gpurun_out/r2_28_pytest_model.log 2>&1; echo "pytest model rc=$?"; tail -3 gpurun_out/r2_28_pytest_model.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_28_$name.json 2> gpurun_out/r2_28_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_28_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), round(d['value']/1e6,3), 'e2e', round(d['e2e']['ms_per_step'],4), 'gemm us/launch', round(d['roofline']['us_per_launch'],1))" || tail -3 gpurun_out/r2_28_$name.err; }
run early X=1
run noearly MAP_B200_EARLY_ADAM=0
run early2 X=1
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --profile-steps 1 --timeline gpurun_out/r2_28_timeline.txt > /dev/null 2> gpurun_out/r2_28_tl.err; head -60 gpurun_out/r2_28_timeline.txt
