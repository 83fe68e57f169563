mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 200 -k "dedup or segment or peer or nce or adamw_rows" > gpurun_out/t39a_k2.log 2>&1; echo "k2 tests rc=$?"; tail -n 5 gpurun_out/t39a_k2.log
timeout 300 python -m pytest tests/test_model_gpu.py tests/test_trainer_gpu.py -m gpu -q -x --timeout 200 > gpurun_out/t39a_model.log 2>&1; echo "model tests rc=$?"; tail -n 3 gpurun_out/t39a_model.log
timeout 200 python scripts/bench_embedding.py --dist uniform > gpurun_out/e39a_c5_uniform.json 2> gpurun_out/e39a_c5.err; echo "emb rc=$?"; cat gpurun_out/e39a_c5_uniform.json
timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --timeline gpurun_out/timeline39a_mfp.txt > gpurun_out/b39a_mfp.json 2> gpurun_out/b39a_mfp.err; echo "bench rc=$?"; head -c 230 gpurun_out/b39a_mfp.json; tail -n 3 gpurun_out/b39a_mfp.err
