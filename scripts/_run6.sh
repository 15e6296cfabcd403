mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/tests_all.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/tests_all.log
timeout 300 python scripts/kbench.py --iters 30 --only cd_512ch_f32,cd+mse_512ch_f32,cd_cfg1_f32,cd_f32,cd_bf16 > gpurun_out/kbench3.log 2>&1; cat gpurun_out/kbench3.log
