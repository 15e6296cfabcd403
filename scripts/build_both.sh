#!/bin/bash
# normal library in place + a timing build (SD_NVCC_EXTRA flags in $1) under build/alt/lib_timing.so
set -e
SD_NVCC_EXTRA="$1" python -m segdistill_b200.build --force | tail -1
mkdir -p build/alt && cp segdistill_b200/libsegdistill_sm100.so build/alt/lib_timing.so
python -m segdistill_b200.build --force | tail -1
