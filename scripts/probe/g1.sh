mkdir -p gpurun_out
timeout 500 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/tests.log
timeout 300 python bench.py --extra --steps 10 --warmup 3 --no-cpu-baseline --no-aten --no-e2e 2> gpurun_out/bench_extra.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d.get('extra') or d
for k in ('cfg2_cgd_4stages_b16_f32_grouped','cfg2_cgd_4stages_b16_f32'): print(k, x[k])"
