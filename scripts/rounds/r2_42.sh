# round 2, call 42: epilogue with pipelined TMEM loads on single-accumulator tiles (build_exp, -DMAP_GEMM_PIPELINED_TMEM_LD) against HEAD
mkdir -p gpurun_out
t() { timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "bf16s" 2>&1 | tail -1; }
b() { timeout 200 python bench.py --steps 150 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],4), round(d['value']/1e6,3), 'gemm us/launch', round(d['roofline']['us_per_launch'],1))"; }
t; b head
cp map_code_b200/libmap_b200.so /tmp/head.so; cp build_exp/libmap_b200.so map_code_b200/libmap_b200.so
t; b exp
cp /tmp/head.so map_code_b200/libmap_b200.so
