# round 2, call 34: validation of HEAD on one GPU: full GPU suite, smoke, default bench, reference arm, ncu launch list of the bench
# command and ncu --set full of the 8 GEMM launches of one step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_34_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_34_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_34_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_34_smoke.log
timeout 600 python bench.py > gpurun_out/r2_34_bench.json 2> gpurun_out/r2_34_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_34_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_34_bench_reference.json 2> gpurun_out/r2_34_bench_reference.err; echo "reference rc=$?"; cut -c1-400 gpurun_out/r2_34_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_34_ncu_launch_list.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-secondary > gpurun_out/r2_34_ncu_list.log 2>&1; echo "ncu list rc=$?"; tail -2 gpurun_out/r2_34_ncu_list.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16s -s 16 -c 8 -o gpurun_out/r2_34_gemm_full -f python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-secondary > gpurun_out/r2_34_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r2_34_ncu_full.log | cut -c1-300
