set -x
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 900 -k "gemm" > gpurun_out/t6_kernels.log 2>&1; echo "kernels rc=$?"; tail -n 3 gpurun_out/t6_kernels.log
python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 900 > gpurun_out/t6_model.log 2>&1; echo "model rc=$?"; tail -n 4 gpurun_out/t6_model.log
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --dump-profile gpurun_out/p6_shapes.txt > gpurun_out/b6_mfp.json 2> gpurun_out/b6_mfp.err; echo "bench rc=$?"; head -c 500 gpurun_out/b6_mfp.json; tail -n 3 gpurun_out/b6_mfp.err
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --task RFD > gpurun_out/b6_rfd.json 2> gpurun_out/b6_rfd.err; echo "bench rfd rc=$?"; head -c 400 gpurun_out/b6_rfd.json
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --optimizer-mode dense_exact > gpurun_out/b6_dense.json 2> gpurun_out/b6_dense.err; echo "bench dense rc=$?"; head -c 400 gpurun_out/b6_dense.json
