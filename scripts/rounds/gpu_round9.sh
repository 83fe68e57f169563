set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/t9_kernels.log 2>&1; echo "kernels rc=$?"; tail -n 5 gpurun_out/t9_kernels.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 300 > gpurun_out/t9_model.log 2>&1; echo "model rc=$?"; tail -n 5 gpurun_out/t9_model.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --dump-profile gpurun_out/p9_shapes.txt > gpurun_out/b9_mfp.json 2> gpurun_out/b9_mfp.err; echo "bench rc=$?"; head -c 300 gpurun_out/b9_mfp.json; tail -n 3 gpurun_out/b9_mfp.err
timeout 600 python scripts/bench_embedding.py > gpurun_out/e9_c5_criteo.json 2> gpurun_out/e9_c5.err; echo "emb rc=$?"; cat gpurun_out/e9_c5_criteo.json; tail -n 3 gpurun_out/e9_c5.err
timeout 600 python scripts/bench_embedding.py --dist uniform > gpurun_out/e9_c5_uniform.json 2>> gpurun_out/e9_c5.err; echo "emb2 rc=$?"; cat gpurun_out/e9_c5_uniform.json
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/plain9.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/ncu9a.log 2>&1; echo "ncu list rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/plain9b.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 44 -c 6 -o gpurun_out/prof_gemm_r1c python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/ncu9b.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
