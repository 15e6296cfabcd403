mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "(fused_bilinear_resize_pixel_mode and dtype0) or (fused_bilinear_resize_matches_reference_resize and dtype0)" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/racecheck.log
