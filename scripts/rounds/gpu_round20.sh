set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/t20_all.log 2>&1; echo "pytest gpu rc=$?"; tail -n 5 gpurun_out/t20_all.log
( time timeout 600 python bench.py ) > gpurun_out/b20_default.json 2> gpurun_out/b20_default.err; echo "bench default rc=$?"; head -c 300 gpurun_out/b20_default.json; tail -n 4 gpurun_out/b20_default.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/b20_ref.json 2> gpurun_out/b20_ref.err; echo "bench ref rc=$?"; head -c 400 gpurun_out/b20_ref.json; tail -n 4 gpurun_out/b20_ref.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 100 --warmup 10 --workload c4 --no-cpu-baseline > gpurun_out/b20_c4_2gpu.json 2> gpurun_out/b20_c4_2gpu.err; echo "bench c4 2gpu rc=$?"; head -c 330 gpurun_out/b20_c4_2gpu.json; grep -v "^\s*$" gpurun_out/b20_c4_2gpu.err | tail -n 5
timeout 300 python bench.py --steps 100 --warmup 10 --workload c4 --no-cpu-baseline > gpurun_out/b20_c4.json 2> gpurun_out/b20_c4.err; echo "bench c4 rc=$?"; head -c 330 gpurun_out/b20_c4.json; tail -n 3 gpurun_out/b20_c4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/b20_mfp_2gpu.json 2> gpurun_out/b20_mfp_2gpu.err; echo "bench 2gpu rc=$?"; head -c 330 gpurun_out/b20_mfp_2gpu.json; grep -v "^\s*$" gpurun_out/b20_mfp_2gpu.err | tail -n 5
