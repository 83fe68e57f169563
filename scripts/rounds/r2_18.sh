# round 2, call 18: xDeepFM / CIN (module path): goldens, CIN vs the reference formulation at the Criteo shape, Trainer loops
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "xdeepfm or cin" > gpurun_out/r2_18_xdeepfm.log 2>&1; echo "model rc=$?"; tail -25 gpurun_out/r2_18_xdeepfm.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -m gpu -x -q -k "xDeepFM or dynamic_mask" > gpurun_out/r2_18_trainer.log 2>&1; echo "trainer rc=$?"; tail -25 gpurun_out/r2_18_trainer.log
