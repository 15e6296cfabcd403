timeout 600 python -m pytest tests -x -q -m gpu -k "corr" 2>&1 | tail -2
timeout 100 python scripts/kbench.py --only corr10_f32,corr10_bf16,corr150_bf16,corr256_512ch_bf16,corr256_512ch_f32 2>&1
