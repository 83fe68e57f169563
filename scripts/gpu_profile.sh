set -x
mkdir -p gpurun_out
python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 900 > gpurun_out/t3_model.log 2>&1; echo "model rc=$?"
tail -n 6 gpurun_out/t3_model.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --dump-profile gpurun_out/p1_shapes.txt > gpurun_out/b2_mfp.json 2> gpurun_out/b2_mfp.err; echo "bench rc=$?"
cat gpurun_out/p1_shapes.txt
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --task RFD > gpurun_out/b2_rfd.json 2> gpurun_out/b2_rfd.err; echo "bench rfd rc=$?"; head -c 700 gpurun_out/b2_rfd.json; tail -n 3 gpurun_out/b2_rfd.err
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --optimizer-mode dense_exact > gpurun_out/b2_mfp_dense.json 2> gpurun_out/b2_mfp_dense.err; echo "bench dense rc=$?"; head -c 400 gpurun_out/b2_mfp_dense.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/ncu1.log 2>&1; echo "ncu list rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 30 -c 4 -o gpurun_out/prof_gemm_r1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
