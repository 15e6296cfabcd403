mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "grouped or groups or grid or two_loss or full_size or near" > gpurun_out/tests_grid.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/tests_grid.log
timeout 100 python scripts/kbench.py --only fused_f32,fused_bf16,cgd10_f32,cgd10_bf16,cd_bf16_grid,cd_f32_grid,cfg2_grouped_f32 2>&1
