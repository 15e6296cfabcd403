mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "corr" > gpurun_out/tests_corr.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|AssertionError: \(" gpurun_out/tests_corr.log | head -30
