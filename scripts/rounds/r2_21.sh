# round 2, call 21: is the GEMM main loop bound by bytes in flight (stage depth) or by L2 bandwidth?  stage-count sensitivity
mkdir -p gpurun_out
LEVELS=S1,S2,S3,S4,B2 timeout 300 python scripts/bench_gemm_bf16s.py > gpurun_out/r2_21_stages_default.txt 2>&1; cat gpurun_out/r2_21_stages_default.txt | grep -v "^ *$"
MAP_B200_STAGES=2 LEVELS=S2,S3,S4,B2 timeout 300 python scripts/bench_gemm_bf16s.py > gpurun_out/r2_21_stages_2.txt 2>&1; cat gpurun_out/r2_21_stages_2.txt
