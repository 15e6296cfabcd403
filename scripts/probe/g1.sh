mkdir -p gpurun_out
timeout 500 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/tests.log
timeout 300 python bench.py --extra --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_extra.log 2> gpurun_out/bench_extra.err; echo "extra rc=$?"; tail -c 400 gpurun_out/bench_extra.err
timeout 300 python scripts/kbench.py > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"; tail -3 gpurun_out/kbench.log
