mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "occupies_the_sms or two_streams" > gpurun_out/tests_grid.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/tests_grid.log
