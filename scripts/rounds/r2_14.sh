# round 2, call 14: whole -m gpu suite at HEAD + ncu of the by-field encoder kernels (MAP_B200_FIELD_ENC=2 runs all four)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_14_pytest_gpu_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_14_pytest_gpu_full.log
export MAP_B200_FIELD_ENC=2
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_14_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"field_enc|head_bwd_fold|field_bucket|nce_fwd" -s 8 -c 5 -o gpurun_out/r2_14_fieldenc python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --profile-steps 1 > gpurun_out/r2_14_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2_14_fieldenc.ncu-rep --page raw --csv > gpurun_out/r2_14_fieldenc_raw.csv 2>/dev/null; echo "export rc=$?"
python scripts/ncu_summary.py gpurun_out/r2_14_fieldenc_raw.csv --json gpurun_out/r2_14_fieldenc_summary.json | cut -c1-400
