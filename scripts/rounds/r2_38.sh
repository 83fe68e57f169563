# round 2, call 38: MFP encoder by field (hybrid) vs dense GEMM + gather at HEAD (the GEMMs are 25 % faster than when hybrid became the default)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" > gpurun_out/r2_38_$name.json 2> gpurun_out/r2_38_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_38_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), round(d['value']/1e6,3), 'e2e', round(d['e2e']['ms_per_step'],4))" || tail -3 gpurun_out/r2_38_$name.err; }
B="timeout 200 python bench.py --steps 150 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1"
run hybrid X=1 $B
run dense MAP_B200_FIELD_ENC=0 $B
run hybrid2 X=1 $B
run dense2 MAP_B200_FIELD_ENC=0 $B
run c4_hybrid X=1 $B --workload c4
run c4_dense MAP_B200_FIELD_ENC=0 $B --workload c4
