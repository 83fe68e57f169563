# round 2, call 43: NCE id sort under the forward pass (MAP_B200_NCE_SORT=early) vs under the backward (default), with the 96-CTA sort grid
mkdir -p gpurun_out
b() { name=$1; shift; env "$@" timeout 100 python bench.py --steps 120 --warmup 10 --no-cpu-baseline --no-secondary --profile-steps 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), round(d['value']/1e6,3))"; }
b late X=1
b early MAP_B200_NCE_SORT=early
MAP_B200_NCE_SORT=early timeout 120 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "mfp" 2>&1 | tail -1
