set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_trainer_gpu.py -m gpu -q --timeout 300 > gpurun_out/t8_model.log 2>&1; echo "model rc=$?"; tail -n 12 gpurun_out/t8_model.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --dump-profile gpurun_out/p8_shapes.txt > gpurun_out/b8_mfp.json 2> gpurun_out/b8_mfp.err; echo "bench rc=$?"; head -c 400 gpurun_out/b8_mfp.json; tail -n 3 gpurun_out/b8_mfp.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --task RFD > gpurun_out/b8_rfd.json 2> gpurun_out/b8_rfd.err; echo "bench rfd rc=$?"; head -c 400 gpurun_out/b8_rfd.json; tail -n 3 gpurun_out/b8_rfd.err
timeout 600 python scripts/bench_embedding.py > gpurun_out/e8_c5_criteo.json 2> gpurun_out/e8_c5.err; echo "emb rc=$?"; cat gpurun_out/e8_c5_criteo.json; tail -n 3 gpurun_out/e8_c5.err
timeout 600 python scripts/bench_embedding.py --dist uniform > gpurun_out/e8_c5_uniform.json 2>> gpurun_out/e8_c5.err; echo "emb2 rc=$?"; cat gpurun_out/e8_c5_uniform.json
