# round 2, call 9 (N GPUs): scaling points with the in-run parity block
N=$1
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N "$@" > gpurun_out/r2_09_${name}_${N}gpu.json 2> gpurun_out/r2_09_${name}_${N}gpu.err; echo "$name rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_09_${name}_${N}gpu.json').read().strip().splitlines()[-1]); print('$name', 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'parity', d.get('parity') and {k: d['parity'][k] for k in ('loss_rel','update_rel','update_rel_worst')}, d['clocks'])" || tail -5 gpurun_out/r2_09_${name}_${N}gpu.err; }
run mfp --steps 100 --warmup 10 --profile-steps 1 --timeline gpurun_out/r2_09_timeline_mfp_${N}gpu.txt
run rfd --task RFD --steps 100 --warmup 10 --profile-steps 1
run c4 --workload c4 --steps 100 --warmup 10 --profile-steps 1
if [ "$N" = "8" ]; then
  run c5_global65536 --workload c5 --batch 8192 --steps 20 --warmup 3 --profile-steps 1
  run c5_rfd_global65536 --workload c5 --task RFD --batch 8192 --steps 20 --warmup 3 --profile-steps 1
fi
