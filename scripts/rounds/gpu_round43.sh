mkdir -p gpurun_out
timeout 150 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --mask-ratio 0.3 > gpurun_out/b43_mfp_mr03.json 2> gpurun_out/b43_mfp_mr03.err; echo "bench mfp 0.3 rc=$?"; head -c 250 gpurun_out/b43_mfp_mr03.json; echo; tail -n 3 gpurun_out/b43_mfp_mr03.err
timeout 150 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --mask-ratio 0.3 --task RFD > gpurun_out/b43_rfd_mr03.json 2> gpurun_out/b43_rfd_mr03.err; echo "bench rfd 0.3 rc=$?"; head -c 250 gpurun_out/b43_rfd_mr03.json; echo; tail -n 3 gpurun_out/b43_rfd_mr03.err
