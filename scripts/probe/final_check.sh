# the contract steps of the round end on one GPU: tests, smoke, both bench arms
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" > gpurun_out/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/tests.log; cat gpurun_out/summary.txt; tail -1 gpurun_out/bench.log | cut -c1-200
