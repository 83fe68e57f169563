set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_trainer_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/t23.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/t23.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --timeline gpurun_out/timeline23_mfp.txt > gpurun_out/b23_mfp.json 2> gpurun_out/b23_mfp.err; echo "bench rc=$?"; head -c 330 gpurun_out/b23_mfp.json; tail -n 3 gpurun_out/b23_mfp.err
MAP_B200_GEMM_GROUP=0 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/b23_mfp_nogroup.json 2> gpurun_out/b23_mfp_nogroup.err; echo "bench nogroup rc=$?"; head -c 330 gpurun_out/b23_mfp_nogroup.json
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --task RFD > gpurun_out/b23_rfd.json 2> gpurun_out/b23_rfd.err; echo "bench rfd rc=$?"; head -c 330 gpurun_out/b23_rfd.json; tail -n 3 gpurun_out/b23_rfd.err
timeout 300 python bench.py --steps 100 --warmup 10 --workload c4 --no-cpu-baseline > gpurun_out/b23_c4.json 2> gpurun_out/b23_c4.err; echo "bench c4 rc=$?"; head -c 330 gpurun_out/b23_c4.json; tail -n 3 gpurun_out/b23_c4.err
