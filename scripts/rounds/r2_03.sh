# round 2, call 3: where does the time of the split-bf16 GEMM go (per-level timing + per-CTA timeline), parity with teacher forcing
mkdir -p gpurun_out; rm -f gpurun_out/parity_fullshape.jsonl
timeout 600 python scripts/bench_gemm_bf16s.py > gpurun_out/r2_03_gemm_levels.txt 2> gpurun_out/r2_03_gemm_levels.err; echo "levels rc=$?"; cat gpurun_out/r2_03_gemm_levels.txt; tail -3 gpurun_out/r2_03_gemm_levels.err
timeout 900 python -m pytest tests/test_fullshape_gpu.py -q -s > gpurun_out/r2_03_fullshape.log 2>&1; echo "fullshape rc=$?"; tail -8 gpurun_out/r2_03_fullshape.log
python - <<'PY'
import json
try:
    for l in open("gpurun_out/parity_fullshape.jsonl"):
        d=json.loads(l)
        print(d["workload"],d["task"],d["optimizer_mode"],"grad %.2e upd %.2e loss %.2e"%(d["max_grad_rel"],d["max_update_rel"],d["max_loss_rel"]), "logits", ["%.1e"%s["logits_rel"] for s in d["steps"]])
        for i,st in enumerate(d["steps"]):
            worst=sorted(st["grad_rel"].items(), key=lambda kv:-kv[1])[:3]
            print("   step",i,"worst grads:", [(k,"%.2e"%v) for k,v in worst], "worst upd", max(st["update_rel"].items(), key=lambda kv: kv[1]))
except Exception as e: print(e)
PY
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_trainer_gpu.py -q > gpurun_out/r2_03_model.log 2>&1; echo "model rc=$?"; tail -12 gpurun_out/r2_03_model.log
