mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/tests.log
for c in 32 64; do echo "pix cols=$c"; SEGDISTILL_PIX_COLS=$c timeout 100 python scripts/kbench.py --only pd_f32,pd_bf16 2>&1; done
