mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "grouped or groups or grid" > gpurun_out/tests_grid.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/tests_grid.log
for g in 1 0; do echo "group_grid=$g"; SEGDISTILL_GROUP_GRID=$g timeout 100 python scripts/kbench.py --only cfg2_grouped_f32,cfg2_separate_f32 2>&1; done
