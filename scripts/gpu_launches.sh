#!/bin/bash
# ncu launch lists (device time per launch) of the fused and the unfused bench step -> gpurun_out/
mkdir -p gpurun_out
for mode in fused unfused; do
  flag=""; [ $mode = unfused ] && flag="--unfused"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kl_rows|scale_grad|mse_|kl_pixels' -s 12 -c 24 --csv \
     --log-file gpurun_out/launches_$mode.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e $flag \
     > gpurun_out/launches_${mode}_run.log 2>&1
  echo "== $mode"; grep -E '"sd::|"void sd::' gpurun_out/launches_$mode.csv | awk -F'","' '{print $5, $NF}' | tail -12
done
