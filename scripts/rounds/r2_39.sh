# round 2, call 39: validation of HEAD on one GPU (dense MFP encoder default for DCNv2 again): full GPU suite, smoke, default bench,
# reference arm, ncu launch list, ncu --set full of the 9 GEMM launches of one step, timeline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_39_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_39_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_39_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_39_smoke.log
timeout 600 python bench.py > gpurun_out/r2_39_bench.json 2> gpurun_out/r2_39_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_39_bench.json
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --profile-steps 1 --timeline gpurun_out/r2_39_timeline.txt > /dev/null 2> gpurun_out/r2_39_tl.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_39_ncu_launch_list.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-secondary > gpurun_out/r2_39_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16s -s 18 -c 9 -o gpurun_out/r2_39_gemm_full -f python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-secondary > gpurun_out/r2_39_ncu_full.log 2>&1; echo "ncu full rc=$?"
