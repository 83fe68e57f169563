# round 2, call 1: parity at the benchmarked shapes (TF32 default + exact-fp32 SIMT for scale), first run of the full C5 step
mkdir -p gpurun_out
rm -f gpurun_out/parity_fullshape.jsonl
timeout 900 python -m pytest tests/test_fullshape_gpu.py -q -s > gpurun_out/r2_01_fullshape.log 2>&1; echo "fullshape rc=$?"
mv gpurun_out/parity_fullshape.jsonl gpurun_out/r2_01_parity_tf32.jsonl
MAP_B200_GEMM=simt timeout 600 python -m pytest tests/test_fullshape_gpu.py -q -s -k "c2 and sparse" > gpurun_out/r2_01_fullshape_simt.log 2>&1; echo "simt rc=$?"
mv gpurun_out/parity_fullshape.jsonl gpurun_out/r2_01_parity_simt.jsonl
python - <<'PY'
import json
for f in ("gpurun_out/r2_01_parity_tf32.jsonl","gpurun_out/r2_01_parity_simt.jsonl"):
    try:
        for l in open(f):
            d=json.loads(l)
            print(f.split('_')[-1], d["workload"],d["task"],d["optimizer_mode"],"grad %.2e upd %.2e loss %.2e"%(d["max_grad_rel"],d["max_update_rel"],d["max_loss_rel"]), "logits", [round(s["logits_rel"],6) for s in d["steps"]])
            worst=sorted(d["steps"][0]["grad_rel"].items(), key=lambda kv:-kv[1])[:6]
            print("   worst grads:", [(k,"%.2e"%v) for k,v in worst])
    except Exception as e: print(f, e)
PY
MAP_B200_BENCH_VERBOSE=1 timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/r2_01_c5_mfp.json 2> gpurun_out/r2_01_c5_mfp.err; echo "c5 rc=$?"; tail -c 1500 gpurun_out/r2_01_c5_mfp.json; tail -5 gpurun_out/r2_01_c5_mfp.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
