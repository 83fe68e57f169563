mkdir -p gpurun_out
# 1. the kernels touched this round first (persistent dedup, pos_seg segment reduce, peer merge), both dedup pipelines
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 300 -k "dedup or segment or peer or nce or adamw_rows" > gpurun_out/t31_k2.log 2>&1; echo "k2 tests rc=$?"; tail -n 5 gpurun_out/t31_k2.log
# 2. the whole GPU suite (what the driver runs)
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/t31_all.log 2>&1; echo "all tests rc=$?"; tail -n 5 gpurun_out/t31_all.log
# 3. bench
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --timeline gpurun_out/timeline31_mfp.txt > gpurun_out/b31_mfp.json 2> gpurun_out/b31_mfp.err; echo "bench rc=$?"; head -c 230 gpurun_out/b31_mfp.json; tail -n 3 gpurun_out/b31_mfp.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --task RFD > gpurun_out/b31_rfd.json 2> gpurun_out/b31_rfd.err; echo "bench rfd rc=$?"; head -c 230 gpurun_out/b31_rfd.json; tail -n 3 gpurun_out/b31_rfd.err
MAP_B200_DEDUP=multi timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/b31_mfp_multi.json 2> gpurun_out/b31_mfp_multi.err; echo "bench multi rc=$?"; head -c 230 gpurun_out/b31_mfp_multi.json
# 4. embedding kernels at C5 scale
timeout 300 python scripts/bench_embedding.py > gpurun_out/e31_c5_criteo.json 2> gpurun_out/e31_c5.err; echo "emb rc=$?"; cat gpurun_out/e31_c5_criteo.json
timeout 300 python scripts/bench_embedding.py --dist uniform > gpurun_out/e31_c5_uniform.json 2>> gpurun_out/e31_c5.err; echo "emb rc=$?"; cat gpurun_out/e31_c5_uniform.json
MAP_B200_DEDUP=multi timeout 300 python scripts/bench_embedding.py --dist uniform > gpurun_out/e31_c5_uniform_multi.json 2>> gpurun_out/e31_c5.err; echo "emb multi rc=$?"; cat gpurun_out/e31_c5_uniform_multi.json
