# round 2, call 17 (1 GPU): the driver's default command + the other single-GPU points at HEAD (C3, C4, full C5 step MFP / RFD)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_17_bench_default.json 2> gpurun_out/r2_17_bench_default.err; echo "default rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_17_bench_reference_arm.json 2> gpurun_out/r2_17_bench_reference_arm.err; echo "reference rc=$?"
run() { name=$1; shift; timeout 900 python bench.py "$@" --no-cpu-baseline --no-secondary > gpurun_out/r2_17_${name}_1gpu.json 2> gpurun_out/r2_17_${name}_1gpu.err; echo "$name rc=$?"; }
run rfd --task RFD --steps 100 --warmup 10
run c4 --workload c4 --steps 100 --warmup 10
run c5_mfp --workload c5 --steps 20 --warmup 3
run c5_rfd --workload c5 --task RFD --steps 20 --warmup 3
for f in default rfd_1gpu c4_1gpu c5_mfp_1gpu c5_rfd_1gpu; do n=gpurun_out/r2_17_$f.json; [ -f gpurun_out/r2_17_bench_$f.json ] && n=gpurun_out/r2_17_bench_$f.json; python -c "
import json; d=json.loads(open('$n').read().strip().splitlines()[-1]); print('$f', 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],4), d['clocks'])"; done
