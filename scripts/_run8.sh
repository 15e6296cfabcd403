mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "cluster or two_losses or dispatcher or full_size or properties" > gpurun_out/tests_cl.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/tests_cl.log
timeout 300 python scripts/kbench.py --iters 30 --only fused_f32,fused_bf16 > gpurun_out/kbench_cl.log 2>&1; cat gpurun_out/kbench_cl.log
