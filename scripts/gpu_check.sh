#!/bin/bash
# One gpurun call: smoke, GPU parity tests, short bench, ncu launch list.  Logs -> gpurun_out/.
# usage: scripts/gpu_check.sh [stage ...]   stages: smoke sanitize tests bench extra launches ncu
set -u
mkdir -p gpurun_out
STAGES=${@:-smoke tests bench launches}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for st in $STAGES; do
  case $st in
    smoke)
      timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt; tail -8 gpurun_out/smoke.log;;
    sanitize)
      timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitize.log 2>&1; echo "sanitize rc=$?" | tee -a gpurun_out/summary.txt; tail -15 gpurun_out/sanitize.log;;
    tests)
      timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/summary.txt; tail -40 gpurun_out/tests.log;;
    bench)
      timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/bench.log;;
    extra)
      timeout 600 python bench.py --steps 20 --warmup 5 --extra --no-cpu-baseline > gpurun_out/bench_extra.log 2>&1; echo "extra rc=$?" | tee -a gpurun_out/summary.txt; tail -2 gpurun_out/bench_extra.log;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches.csv \
         python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/launches_run.log 2>&1; echo "launches rc=$?" | tee -a gpurun_out/summary.txt; tail -12 gpurun_out/launches.csv;;
    ncu)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:kl_rows -s 3 -c 2 -o gpurun_out/prof_bench -f \
         python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_run.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/ncu_run.log;;
    ncu_k)
      # one capture per kernel family through the C ABI (scripts/kbench.py): CD (register kernel), CGD (stream), PD (pixels)
      for c in cd_f32 cgd10_f32 pd_f32 cd_bf16; do
        timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kl_rows|kl_pixels' -s 6 -c 1 -o gpurun_out/prof_$c -f \
           python scripts/kbench.py --iters 3 --only $c > gpurun_out/ncu_$c.log 2>&1; echo "ncu_k $c rc=$?" | tee -a gpurun_out/summary.txt
      done;;
    kbench)
      timeout 600 python scripts/kbench.py --iters 50 > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?" | tee -a gpurun_out/summary.txt; cat gpurun_out/kbench.log;;
  esac
done
cat gpurun_out/summary.txt
