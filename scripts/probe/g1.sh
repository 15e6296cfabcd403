mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu --timeout 60 --timeout-method=thread > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
