mkdir -p gpurun_out
for c in ${CASES:-fused_f32}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kl_rows|kl_pixels' -s 6 -c 1 -o gpurun_out/prof_$c -f python scripts/kbench.py --iters 3 --only $c > gpurun_out/ncu_$c.log 2>&1; echo "rc=$?"
done
