# round 2, call 2: first run of the split-bf16 GEMM (unit tests vs fp64), then the step on top of it
mkdir -p gpurun_out; rm -f gpurun_out/gemm_bf16s_errors.jsonl gpurun_out/parity_fullshape.jsonl
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "split_planes or bf16s" > gpurun_out/r2_02_bf16s.log 2>&1; rc=$?; echo "bf16s rc=$rc"; tail -15 gpurun_out/r2_02_bf16s.log
python - <<'PY'
import json
try:
    rows=[json.loads(l) for l in open("gpurun_out/gemm_bf16s_errors.jsonl")]
    for t in (3,6):
        r=[x for x in rows if x["terms"]==t]
        print("terms",t,"n",len(r),"max rel %.2e"%max(x["rel"] for x in r),"median %.2e"%sorted(x["rel"] for x in r)[len(r)//2])
        for x in sorted(r,key=lambda x:-x["rel"])[:4]: print("   ",x)
except Exception as e: print("no error log", e)
PY
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "not bf16s" > gpurun_out/r2_02_kernels.log 2>&1; echo "kernels rc=$?"; tail -3 gpurun_out/r2_02_kernels.log
timeout 900 python -m pytest tests/test_fullshape_gpu.py -q -s > gpurun_out/r2_02_fullshape.log 2>&1; echo "fullshape rc=$?"; tail -8 gpurun_out/r2_02_fullshape.log
python - <<'PY'
import json
try:
    for l in open("gpurun_out/parity_fullshape.jsonl"):
        d=json.loads(l)
        print(d["workload"],d["task"],d["optimizer_mode"],"grad %.2e upd %.2e loss %.2e"%(d["max_grad_rel"],d["max_update_rel"],d["max_loss_rel"]), "logits", ["%.1e"%s["logits_rel"] for s in d["steps"]])
        for i,st in enumerate(d["steps"]):
            worst=sorted(st["grad_rel"].items(), key=lambda kv:-kv[1])[:4]
            print("   step",i,"worst grads:", [(k,"%.2e"%v) for k,v in worst])
except Exception as e: print(e)
PY
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_trainer_gpu.py -q -x > gpurun_out/r2_02_model.log 2>&1; echo "model rc=$?"; tail -12 gpurun_out/r2_02_model.log
timeout 600 python bench.py --steps 100 --warmup 10 --profile-steps 2 > gpurun_out/r2_02_bench_mfp.json 2> gpurun_out/r2_02_bench_mfp.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r2_02_bench_mfp.json; tail -3 gpurun_out/r2_02_bench_mfp.err
python -c "import json;d=json.loads(open('gpurun_out/r2_02_bench_mfp.json').read().strip().splitlines()[-1]);print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','roofline','gpu_eager_reference','secondary','cpu_baseline','impl_detail')},indent=1))"
