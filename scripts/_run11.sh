mkdir -p gpurun_out
timeout 300 python scripts/kbench.py --iters 20 --only up4_cgd10_2x150x128_f32,up4_cgd10_16x150x128_f32,up8_cd_16x150x64_f32,up4_cgd10_16x150x128_bf16 > gpurun_out/kbench_up.log 2>&1; cat gpurun_out/kbench_up.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kl_rows_up' -s 8 -c 2 -o gpurun_out/prof_up4 -f python scripts/kbench.py --iters 3 --only up4_cgd10_16x150x128_f32 > gpurun_out/ncu_up4.log 2>&1; echo "rc=$?"
