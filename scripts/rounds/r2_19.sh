# round 2, call 19: full GPU suite + smoke + default bench at HEAD after the container was re-created (fresh build)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_19_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_19_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_19_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_19_smoke.log
timeout 600 python bench.py > gpurun_out/r2_19_bench.json 2> gpurun_out/r2_19_bench.err; echo "bench rc=$?"; cat gpurun_out/r2_19_bench.json | cut -c1-1500
