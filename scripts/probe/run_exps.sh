# swaps experiment builds of the library in and runs the cycle-counter script on each
cp segdistill_b200/libsegdistill_sm100.so /tmp/lib_orig.so
for f in scripts/probe/libs/lib_exp*.so; do
  echo "=== $f"; cp $f segdistill_b200/libsegdistill_sm100.so
  timeout 120 python scripts/cluster_timing.py 2>&1 | grep -E "pk:|gr:|st:merge|st:wait_xch"
done
cp /tmp/lib_orig.so segdistill_b200/libsegdistill_sm100.so
