for i in 1 2 3; do timeout 100 python scripts/kbench.py --only fused_f32,cfg2_grouped_f32 2>&1; done
