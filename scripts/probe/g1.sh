mkdir -p gpurun_out
for k in 3,96,32,64 3,400,100,200 3,1000,200,400 2,2000,500,500; do
echo "knobs=$k"; SEGDISTILL_GRID_KNOBS=$k timeout 100 python scripts/kbench.py --only fused_f32,fused_bf16,cgd10_bf16 2>&1 | tee -a gpurun_out/kbench_grid.log; done
