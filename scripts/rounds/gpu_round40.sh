mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/t40_all.log 2>&1; echo "all tests rc=$?"; tail -n 3 gpurun_out/t40_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke40.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke40.log
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/b40_mfp.json 2> gpurun_out/b40_mfp.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/b40_mfp.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['traffic'], d['roofline']['frac'], d['gpu_launches'])"; tail -n 3 gpurun_out/b40_mfp.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/b40_ref.json 2> gpurun_out/b40_ref.err; echo "ref rc=$?"; head -c 300 gpurun_out/b40_ref.json
