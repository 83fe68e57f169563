# round 2, call 35: head level of the hybrid MFP step as ONE grouped launch (CROSS_BWD + RELUMASK combination was missing): tests, bench,
# ncu --set full of the 8 GEMM launches of one step, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_fullshape_gpu.py -m gpu -x -q -k "bf16s or golden or fullshape or benchmarked or peer" > gpurun_out/r2_35_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_35_pytest.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_35_$name.json 2> gpurun_out/r2_35_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_35_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), round(d['value']/1e6,3), 'e2e', round(d['e2e']['ms_per_step'],4), 'gemm us/launch', round(d['roofline']['us_per_launch'],1), d['roofline']['launches_per_step'])" || tail -3 gpurun_out/r2_35_$name.err; }
run head1 X=1
run rfd X=1 
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --profile-steps 1 --timeline gpurun_out/r2_35_timeline.txt > /dev/null 2> gpurun_out/r2_35_tl.err; grep -n "gemm\|step span\|field_enc_fwd\|nce_fwd" gpurun_out/r2_35_timeline.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16s -s 16 -c 8 -o gpurun_out/r2_35_gemm_full -f python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-secondary > gpurun_out/r2_35_ncu_full.log 2>&1; echo "ncu full rc=$?"
