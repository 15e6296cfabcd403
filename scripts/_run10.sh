mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/launches_run.log 2>&1; echo "launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kl_rows_cluster -s 3 -c 1 -o gpurun_out/prof_bench -f python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_run.log 2>&1; echo "ncu rc=$?"
