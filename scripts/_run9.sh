mkdir -p gpurun_out
timeout 300 python scripts/kbench.py --iters 30 --only fused_f32,fused_bf16 > gpurun_out/kbench_cl.log 2>&1; cat gpurun_out/kbench_cl.log
CASES=fused_f32 bash scripts/_run7.sh
