mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_trainer_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/t28.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t28.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --timeline gpurun_out/timeline28_mfp.txt > gpurun_out/b28_mfp.json 2> gpurun_out/b28_mfp.err; echo "bench rc=$?"; head -c 230 gpurun_out/b28_mfp.json; tail -n 3 gpurun_out/b28_mfp.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --task RFD --timeline gpurun_out/timeline28_rfd.txt > gpurun_out/b28_rfd.json 2> gpurun_out/b28_rfd.err; echo "bench rfd rc=$?"; head -c 230 gpurun_out/b28_rfd.json; tail -n 3 gpurun_out/b28_rfd.err
