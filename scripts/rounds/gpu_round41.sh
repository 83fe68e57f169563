mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 4 --steps 100 --warmup 10 > gpurun_out/b41_mfp_4gpu.json 2> gpurun_out/b41_mfp_4gpu.err; echo "bench 4 rc=$?"; head -c 250 gpurun_out/b41_mfp_4gpu.json; echo; grep -v "^\s*$" gpurun_out/b41_mfp_4gpu.err | grep -iv "OMP_NUM\|\*\*\*\*" | tail -n 4
