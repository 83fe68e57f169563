# round 2, call 15: by-field kernels after the fold / dgrad / wgrad fixes: kernel parity + A/B of the three encoder modes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -x -q -k "field_enc or golden" > gpurun_out/r2_15_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_15_tests.log
for fe in 1 2; do
  MAP_B200_FIELD_ENC=$fe timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --timeline gpurun_out/r2_15_timeline_fe$fe.txt --dump-profile gpurun_out/r2_15_profile_fe$fe.txt > gpurun_out/r2_15_bench_fe$fe.json 2> gpurun_out/r2_15_bench_fe$fe.err; echo "bench fe=$fe rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/r2_15_bench_fe$fe.json').read().strip().splitlines()[-1]); print('fe=$fe value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))
for k in d['kernels'][:12]:
    if 'field' in k['kernel'] or 'fold' in k['kernel']: print('   ', k['kernel'], round(k['us_per_step'],1), k.get('gbs') and round(k['gbs']))" || tail -5 gpurun_out/r2_15_bench_fe$fe.err
done
