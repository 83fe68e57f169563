# round 2, call 6 (2 GPUs): sharded step on the split-bf16 backend: 2-GPU parity tests at HEAD, bench N=2 with the in-run parity block
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "full_softmax or shared_noise or feat_count or gather_rows" > gpurun_out/r2_06_newkernels.log 2>&1; echo "newkernels rc=$?"; tail -6 gpurun_out/r2_06_newkernels.log
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -q > gpurun_out/r2_06_dist.log 2>&1; echo "dist rc=$?"; tail -8 gpurun_out/r2_06_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 --profile-steps 1 --timeline gpurun_out/r2_06_timeline_2gpu.txt > gpurun_out/r2_06_bench_2gpu.json 2> gpurun_out/r2_06_bench_2gpu.err; echo "bench2 rc=$?"; tail -5 gpurun_out/r2_06_bench_2gpu.err
python -c "import json;d=json.loads(open('gpurun_out/r2_06_bench_2gpu.json').read().strip().splitlines()[-1]);print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','parity')},indent=1))"
cat gpurun_out/r2_06_timeline_2gpu.txt | head -80
