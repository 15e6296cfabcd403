mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kl_rows_up_grad' -s 4 -c 1 -o gpurun_out/prof_up4g -f python scripts/kbench.py --iters 3 --only up4_cgd10_16x150x128_f32 > gpurun_out/ncu_up4g.log 2>&1; echo "rc=$?"
