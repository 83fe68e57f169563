mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q -x --timeout 90 > gpurun_out/t46_all.log 2>&1; echo "all tests rc=$?"; tail -n 2 gpurun_out/t46_all.log
