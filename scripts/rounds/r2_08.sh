# round 2, call 8 (1 GPU): the driver's commands + the other configurations + ncu evidence
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_08_bench_mfp_default.json 2> gpurun_out/r2_08_bench_mfp_default.err; echo "default rc=$?"; tail -2 gpurun_out/r2_08_bench_mfp_default.err
python -c "import json;d=json.loads(open('gpurun_out/r2_08_bench_mfp_default.json').read().strip().splitlines()[-1]);print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','config','gpu_eager_reference','secondary','cpu_baseline','clocks')},indent=1)); print(d['roofline'])"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_08_bench_reference.json 2> gpurun_out/r2_08_bench_reference.err; echo "reference rc=$?"; head -c 900 gpurun_out/r2_08_bench_reference.json; echo
timeout 600 python bench.py --task RFD --steps 200 --warmup 20 --no-cpu-baseline --no-secondary > gpurun_out/r2_08_bench_rfd.json 2> gpurun_out/r2_08_bench_rfd.err; echo "rfd rc=$?"; head -c 300 gpurun_out/r2_08_bench_rfd.json; echo
timeout 600 python bench.py --workload c4 --steps 200 --warmup 20 --no-cpu-baseline --no-secondary > gpurun_out/r2_08_bench_c4.json 2> gpurun_out/r2_08_bench_c4.err; echo "c4 rc=$?"; head -c 300 gpurun_out/r2_08_bench_c4.json; echo
timeout 900 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_08_bench_c5_mfp.json 2> gpurun_out/r2_08_bench_c5_mfp.err; echo "c5 mfp rc=$?"; head -c 300 gpurun_out/r2_08_bench_c5_mfp.json; echo; tail -2 gpurun_out/r2_08_bench_c5_mfp.err
timeout 900 python bench.py --workload c5 --task RFD --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_08_bench_c5_rfd.json 2> gpurun_out/r2_08_bench_c5_rfd.err; echo "c5 rfd rc=$?"; head -c 300 gpurun_out/r2_08_bench_c5_rfd.json; echo; tail -2 gpurun_out/r2_08_bench_c5_rfd.err
timeout 600 python scripts/bench_embedding.py --dist uniform > gpurun_out/r2_08_embedding_c5_uniform.json 2> gpurun_out/r2_08_embedding.err; echo "emb rc=$?"; tail -c 1200 gpurun_out/r2_08_embedding_c5_uniform.json
# ---- ncu (each only after the same command exited 0 without it)
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-graph --profile-steps 1"
$CMD > gpurun_out/r2_08_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_08_ncu_launch_list.csv $CMD > gpurun_out/r2_08_ncu1.log 2>&1; echo "ncu list rc=$?"
$CMD > gpurun_out/r2_08_plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16s -s 18 -c 9 -o gpurun_out/r2_08_prof_gemm_bf16s $CMD > gpurun_out/r2_08_ncu2.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu -i gpurun_out/r2_08_prof_gemm_bf16s.ncu-rep --page raw --csv > gpurun_out/r2_08_ncu_full_gemm_bf16s_raw.csv 2>/dev/null; echo "export rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nce_fwd|dedup_persistent|segment_reduce|emb_gather|adamw_sparse" -s 10 -c 8 -o gpurun_out/r2_08_prof_tables $CMD > gpurun_out/r2_08_ncu3.log 2>&1; echo "ncu tables rc=$?"
timeout 600 ncu -i gpurun_out/r2_08_prof_tables.ncu-rep --page raw --csv > gpurun_out/r2_08_ncu_full_tables_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -20
