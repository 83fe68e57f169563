mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 300 -k "dedup or segment or peer" > gpurun_out/t32_k2.log 2>&1; echo "k2 tests rc=$?"; tail -n 3 gpurun_out/t32_k2.log
timeout 300 python scripts/bench_dedup.py > gpurun_out/d32_single.txt 2> gpurun_out/d32.err; echo "dedup single rc=$?"; cat gpurun_out/d32_single.txt
MAP_B200_DEDUP=multi timeout 300 python scripts/bench_dedup.py > gpurun_out/d32_multi.txt 2>> gpurun_out/d32.err; echo "dedup multi rc=$?"; cat gpurun_out/d32_multi.txt
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --timeline gpurun_out/timeline32_mfp.txt > gpurun_out/b32_mfp.json 2> gpurun_out/b32_mfp.err; echo "bench rc=$?"; head -c 230 gpurun_out/b32_mfp.json; tail -n 3 gpurun_out/b32_mfp.err
MAP_B200_DEDUP=multi timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/b32_mfp_multi.json 2> gpurun_out/b32_mfp_multi.err; echo "bench multi rc=$?"; head -c 230 gpurun_out/b32_mfp_multi.json
