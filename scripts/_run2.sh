mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "cluster or two_losses or dispatcher" > gpurun_out/tests_cluster.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/tests_cluster.log
timeout 300 python scripts/kbench.py --iters 30 --only cgd10_f32,cgd10_bf16,fused_f32,fused_bf16,cgd10_f32_stream,fused_f32_stream > gpurun_out/kbench2.log 2>&1; cat gpurun_out/kbench2.log
