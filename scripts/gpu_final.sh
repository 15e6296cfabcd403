#!/bin/bash
# End-of-round record: smoke, GPU tests, default bench, reference arm, bench --extra, kernel table, launch list -> gpurun_out/
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-200
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --extra --no-cpu-baseline > gpurun_out/bench_extra.log 2>&1; echo "extra rc=$?"
timeout 600 python scripts/kbench.py --iters 50 > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/launches_run.log 2>&1; echo "launches rc=$?"
