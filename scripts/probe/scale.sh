# usage: scale.sh N   (on a box with N GPUs)
N=$1
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; echo rc=$?
tail -1 gpurun_out/scale_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N',d['n_gpus'],'value',round(d['value'],1),'ms_per_step',round(d['ms_per_step'],5),'kernel',d['kernel_ms']['dominant_kernel'],'e2e',round(d['e2e']['value'],1),'clocks',d['clocks'])"
