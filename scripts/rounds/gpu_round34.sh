mkdir -p gpurun_out
for mode in single multi; do
  MAP_B200_DEDUP=$mode timeout 170 compute-sanitizer --tool synccheck python scripts/sanitize_multi.py > gpurun_out/san34_sync_$mode.log 2>&1; echo "synccheck $mode rc=$?"; tail -n 4 gpurun_out/san34_sync_$mode.log
  MAP_B200_DEDUP=$mode timeout 170 compute-sanitizer --tool racecheck python scripts/sanitize_multi.py > gpurun_out/san34_race_$mode.log 2>&1; echo "racecheck $mode rc=$?"; tail -n 4 gpurun_out/san34_race_$mode.log
done
# does the multi-launch pipeline hang in a small fused step? (trainer test, 60 s)
MAP_B200_DEDUP=multi timeout 60 python -m pytest tests/test_trainer_gpu.py -m gpu -q -x -k "async" > gpurun_out/t34_multi.log 2>&1; echo "trainer multi rc=$?"; tail -n 3 gpurun_out/t34_multi.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/t34_all.log 2>&1; echo "all tests rc=$?"; tail -n 3 gpurun_out/t34_all.log
timeout 200 python scripts/bench_dedup.py > gpurun_out/d34_single.txt 2> gpurun_out/d34.err; echo "dedup single rc=$?"; cat gpurun_out/d34_single.txt
timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --timeline gpurun_out/timeline34_mfp.txt > gpurun_out/b34_mfp.json 2> gpurun_out/b34_mfp.err; echo "bench rc=$?"; head -c 230 gpurun_out/b34_mfp.json; tail -n 3 gpurun_out/b34_mfp.err
