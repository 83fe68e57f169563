# round 2, call 33 (8 GPUs): headline at 8 GPUs: peer all-reduce vs NCCL, 72 GEMM clusters
N=$1
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N "$@" > gpurun_out/r2_33_${name}_${N}gpu.json 2> gpurun_out/r2_33_${name}_${N}gpu.err; echo "$name rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_33_${name}_${N}gpu.json').read().strip().splitlines()[-1]); print('$name', 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'parity', d.get('parity') and {k: d['parity'][k] for k in ('loss_rel','update_rel','update_rel_worst')})" || tail -5 gpurun_out/r2_33_${name}_${N}gpu.err; }
run mfp --steps 60 --warmup 10 --profile-steps 1 --no-cpu-baseline --timeline gpurun_out/r2_33_timeline_mfp_${N}gpu.txt
MAP_B200_GRAD_AR=nccl run mfp_nccl --steps 60 --warmup 10 --profile-steps 1 --no-cpu-baseline
MAP_B200_GEMM_CLUSTERS=72 run mfp_c72 --steps 60 --warmup 10 --profile-steps 1 --no-cpu-baseline
