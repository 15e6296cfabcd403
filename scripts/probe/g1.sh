mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/tests.log
timeout 100 python scripts/kbench.py --only cd_512ch_f32,cd+mse_512ch_f32,cd_cfg1_f32,cd_f32,cd_bf16 2>&1
