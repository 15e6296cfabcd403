# swaps experiment builds of the library in and runs a command on each:  run_libs.sh <python args...>
cp segdistill_b200/libsegdistill_sm100.so /tmp/lib_orig.so
for f in scripts/probe/libs/*.so; do
  echo "=== $f"; cp $f segdistill_b200/libsegdistill_sm100.so
  timeout 200 python "$@" 2>&1 | tail -8
done
cp /tmp/lib_orig.so segdistill_b200/libsegdistill_sm100.so
