#!/bin/bash
# End-of-round record on the GPU box: tests, smoke, both bench arms, --extra, kernel table, ncu launch list and one
# full ncu capture of the dominant kernel inside the bench command.  Everything lands in gpurun_out/; every step is
# bounded (a hung step must not take the others with it).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" > gpurun_out/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/summary.txt
timeout 400 python bench.py --extra > gpurun_out/bench_extra.log 2> gpurun_out/bench_extra.err; echo "extra rc=$?" >> gpurun_out/summary.txt
timeout 300 python scripts/kbench.py > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?" >> gpurun_out/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/launches_run.log 2>&1; echo "launches rc=$?" >> gpurun_out/summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kl_rows_grid -s 3 -c 1 -f -o gpurun_out/prof_bench_r02 \
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
tail -2 gpurun_out/tests.log; cat gpurun_out/summary.txt
tail -1 gpurun_out/bench.log | cut -c1-300
