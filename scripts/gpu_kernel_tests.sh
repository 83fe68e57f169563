set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "not tcgen05" --timeout 900 > gpurun_out/t1_base.log 2>&1; echo "base rc=$?"
for combo in False-False False-True True-True True-False; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "tcgen05_tf32 and $combo" > gpurun_out/t1_tc_$combo.log 2>&1; echo "tc $combo rc=$?"
done
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "tcgen05_epilogues or tcgen05_strided" > gpurun_out/t1_tc_epi.log 2>&1; echo "tc epi rc=$?"
for f in gpurun_out/t1_*.log; do echo "== $f"; tail -n 3 $f; done
