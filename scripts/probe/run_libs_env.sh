cp segdistill_b200/libsegdistill_sm100.so /tmp/lib_orig.so
for f in scripts/probe/libs/*.so; do
  echo "=== $f"; cp $f segdistill_b200/libsegdistill_sm100.so
  SD_B=16 timeout 100 python scripts/cluster_timing.py bf16 2>&1 | grep -E "pk:rows|pk:wait|pk:total|gr:"
done
cp /tmp/lib_orig.so segdistill_b200/libsegdistill_sm100.so
