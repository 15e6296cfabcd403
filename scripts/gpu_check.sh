#!/bin/bash
# One GPU visit: build check, smoke, the GPU test-suite, a short bench.  Usage (from the repo root, on the B200 box):
#   bash scripts/gpu_check.sh [pytest -k expression]
set -o pipefail
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -15
if [ -n "$1" ]; then
  python -m pytest tests -q -m gpu -x -k "$1" 2>&1 | tail -25
else
  python -m pytest tests -q -m gpu -x 2>&1 | tail -25
fi
python bench.py --steps 50 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench_quick.json
