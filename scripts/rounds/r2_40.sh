# round 2, call 40 (2 GPUs): validation of HEAD: the 5 multi-GPU tests, headline and C4 with the in-run parity block, timeline
N=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q > gpurun_out/r2_40_dist_gpu_2gpu.log 2>&1; echo "dist tests rc=$?"; tail -3 gpurun_out/r2_40_dist_gpu_2gpu.log
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N "$@" > gpurun_out/r2_40_${name}_${N}gpu.json 2> gpurun_out/r2_40_${name}_${N}gpu.err; echo "$name rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_40_${name}_${N}gpu.json').read().strip().splitlines()[-1]); print('$name', 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'parity', d.get('parity') and {k: d['parity'][k] for k in ('loss_rel','update_rel','update_rel_worst')})" || tail -5 gpurun_out/r2_40_${name}_${N}gpu.err; }
run mfp --steps 100 --warmup 10 --profile-steps 1 --no-cpu-baseline --timeline gpurun_out/r2_40_timeline_mfp_${N}gpu.txt
run rfd --task RFD --steps 100 --warmup 10 --profile-steps 1 --no-cpu-baseline
