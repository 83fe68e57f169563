mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 200 -k "dedup or segment or peer or nce or adamw_rows" > gpurun_out/t39b_k2.log 2>&1; echo "k2 tests rc=$?"; tail -n 5 gpurun_out/t39b_k2.log
timeout 300 python -m pytest tests/test_model_gpu.py tests/test_trainer_gpu.py -m gpu -q -x --timeout 200 > gpurun_out/t39b_model.log 2>&1; echo "model tests rc=$?"; tail -n 3 gpurun_out/t39b_model.log
timeout 200 python scripts/bench_embedding.py --dist uniform > gpurun_out/e39b_c5_uniform.json 2> gpurun_out/e39b_c5.err; echo "emb rc=$?"; cat gpurun_out/e39b_c5_uniform.json
