mkdir -p gpurun_out
timeout 1100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 200 -k "not full_size and not full_batch and not b16 and not properties and not cfg3 and not occupies and not two_streams" > gpurun_out/memcheck_all.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/memcheck_all.log
