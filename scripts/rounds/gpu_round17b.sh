set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 50 --warmup 5 --timeline gpurun_out/timeline17_mfp_2gpu.txt > gpurun_out/b17_mfp_2gpu.json 2> gpurun_out/b17_mfp_2gpu.err; echo "bench2 rc=$?"; head -c 400 gpurun_out/b17_mfp_2gpu.json; grep -v "^\s*$" gpurun_out/b17_mfp_2gpu.err | grep -v "frame #" | tail -n 15
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 50 --warmup 5 --task RFD > gpurun_out/b17_rfd_2gpu.json 2> gpurun_out/b17_rfd_2gpu.err; echo "bench2 rfd rc=$?"; head -c 400 gpurun_out/b17_rfd_2gpu.json; tail -n 5 gpurun_out/b17_rfd_2gpu.err
