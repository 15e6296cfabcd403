mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/tests.log
timeout 100 python scripts/kbench.py --only fused_f32,fused_bf16,cgd10_f32,cgd10_bf16,cfg2_grouped_f32,cgd150_f32 2>&1
