# N = 2 check of the log flush (dist.DeferredLogs.flush_start): in-place AVG all-reduce + one read-back (default) against
# the reduce-a-copy form (SD_LOG_FLUSH_COPY=1), at the driver's K = 20 and at K = 200; per-rank times in the line.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "deferred_logs or dispatcher_batches" 2>&1 | tail -2
run() { # name, env, port, steps
  env $2 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps $4 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/n2_$1.json 2> gpurun_out/n2_$1.err; echo "$1 rc=$?"
}
run new20 SD_X=0 29511 20
run old20 SD_LOG_FLUSH_COPY=1 29512 20
run new200 SD_X=0 29513 200
for f in new20 old20 new200; do grep '^{' gpurun_out/n2_$f.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$f', d['value'], d['ms_per_step'], d.get('log_flush'), d.get('per_rank'))"; done
