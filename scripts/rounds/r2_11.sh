# round 2, call 11: hybrid by-field encoder (forward + wgrad by field, dense dgrad): parity + A/B bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_fullshape_gpu.py tests/test_trainer_gpu.py -m gpu -x -q > gpurun_out/r2_11_model.log 2>&1; echo "model rc=$?"; tail -5 gpurun_out/r2_11_model.log
for fe in 1 2; do
  MAP_B200_FIELD_ENC=$fe timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --timeline gpurun_out/r2_11_timeline_fe$fe.txt --dump-profile gpurun_out/r2_11_profile_fe$fe.txt > gpurun_out/r2_11_bench_fe$fe.json 2> gpurun_out/r2_11_bench_fe$fe.err; echo "bench fe=$fe rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/r2_11_bench_fe$fe.json').read().strip().splitlines()[-1]); print('fe=$fe value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],4), d['roofline']['launches_per_step'], round(d['roofline']['us_per_launch'],1), d['clocks'])
for k in d['kernels'][:10]: print('   ', k['kernel'], round(k['us_per_step'],1), k.get('gbs') and round(k['gbs']), k.get('tflops') and round(k['tflops'],1))" || tail -5 gpurun_out/r2_11_bench_fe$fe.err
done
