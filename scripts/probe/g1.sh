for k in 3,96,32,64 19,96,32,64; do echo "knobs=$k"; SEGDISTILL_GRID_KNOBS=$k timeout 100 python scripts/kbench.py --only fused_f32,cgd10_f32,fused_bf16 2>&1; done
