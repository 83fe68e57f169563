mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/t42_all.log 2>&1; echo "all tests rc=$?"; tail -n 12 gpurun_out/t42_all.log
