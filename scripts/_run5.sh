mkdir -p gpurun_out
timeout 300 python scripts/kbench.py --iters 30 --only corr10_f32,corr10_bf16,corr150_bf16,corr256_512ch_bf16,corr256_512ch_f32 > gpurun_out/kbench_corr.log 2>&1; cat gpurun_out/kbench_corr.log
for c in corr10_bf16 corr256_512ch_bf16; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'corr_' -s 15 -c 3 -o gpurun_out/prof_$c -f python scripts/kbench.py --iters 3 --only $c > gpurun_out/ncu_$c.log 2>&1; echo "rc=$?"
done
