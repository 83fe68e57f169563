mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/t39_all.log 2>&1; echo "all tests rc=$?"; tail -n 3 gpurun_out/t39_all.log
timeout 400 python bench.py > gpurun_out/b39_default.json 2> gpurun_out/b39_default.err; echo "bench default rc=$?"; head -c 400 gpurun_out/b39_default.json; echo; tail -n 3 gpurun_out/b39_default.err
timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --task RFD > gpurun_out/b39_rfd.json 2> gpurun_out/b39_rfd.err; echo "bench rfd rc=$?"; head -c 250 gpurun_out/b39_rfd.json; echo
timeout 200 python scripts/bench_embedding.py > gpurun_out/e39_c5_criteo.json 2> gpurun_out/e39_c5.err; echo "emb rc=$?"; cat gpurun_out/e39_c5_criteo.json
timeout 200 python scripts/bench_embedding.py --dist uniform > gpurun_out/e39_c5_uniform.json 2>> gpurun_out/e39_c5.err; echo "emb rc=$?"; cat gpurun_out/e39_c5_uniform.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1f.csv $CMD > gpurun_out/ncu39a.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 45 -c 9 -o gpurun_out/prof_gemm_r1f $CMD > gpurun_out/ncu39b.log 2>&1; echo "ncu gemm rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dedup_persistent|segment_reduce" -s 8 -c 4 -o gpurun_out/prof_k2_r1f $CMD > gpurun_out/ncu39c.log 2>&1; echo "ncu k2 rc=$?"
ls -la gpurun_out | tail -n 12
