mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/t29_dist.log 2>&1; echo "dist tests rc=$?"; tail -n 3 gpurun_out/t29_dist.log
for N in 2 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N bench.py --gpus $N --steps 100 --warmup 10 --timeline gpurun_out/timeline29_mfp_${N}gpu.txt > gpurun_out/b29_mfp_${N}gpu.json 2> gpurun_out/b29_mfp_${N}gpu.err; echo "bench $N rc=$?"; head -c 250 gpurun_out/b29_mfp_${N}gpu.json; grep -v "^\s*$" gpurun_out/b29_mfp_${N}gpu.err | grep -iv "OMP_NUM\|\*\*\*\*" | tail -n 4
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29640 bench.py --gpus 4 --steps 100 --warmup 10 --task RFD > gpurun_out/b29_rfd_4gpu.json 2> gpurun_out/b29_rfd_4gpu.err; echo "bench rfd 4 rc=$?"; head -c 250 gpurun_out/b29_rfd_4gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --impl reference --steps 5 --warmup 1 > gpurun_out/b29_ref_4gpu.json 2> gpurun_out/b29_ref_4gpu.err; echo "bench ref 4 rc=$?"; head -c 250 gpurun_out/b29_ref_4gpu.json
