set -x
mkdir -p gpurun_out
python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 900 > gpurun_out/t2_model.log 2>&1; echo "model rc=$?"
tail -n 25 gpurun_out/t2_model.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t2_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 5 gpurun_out/t2_smoke.log
python bench.py --steps 50 --warmup 5 > gpurun_out/b1_mfp.json 2> gpurun_out/b1_mfp.err; echo "bench rc=$?"; tail -n 5 gpurun_out/b1_mfp.err; cat gpurun_out/b1_mfp.json
