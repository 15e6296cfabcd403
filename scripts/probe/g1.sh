mkdir -p gpurun_out
for f in 1 0; do echo "fine=$f fused"; SEGDISTILL_GRID_FINE=$f timeout 100 python scripts/grid_timing.py 2>&1 | tee -a gpurun_out/grid_timing.log; done
