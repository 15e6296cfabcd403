#!/usr/bin/env python
"""bench.py - distillation-loss fwd+bwd throughput on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[4], weak scaling): per GPU logits 16x150x128x128 fp32
(= global B=128 on 8 GPUs), one step = CGDLoss(g=10, tau=2, alpha=3) + CDLoss forward AND
backward on that batch through the reference-facing plugin API: the `DistillationLoss`
dispatcher with two `distillation` entries hooking the same logits tensor (`decode_head` and
`decode_head.linear_pred` return the same tensor).  `value` = Mpixel/s with the maps resident in
HBM; `e2e` = same from pinned HOST buffers (H2D of S and T and D2H of the loss scalars inside the
timed region).  Inputs
(2 x 157 MB per GPU) exceed the 126 MB L2, so no L2 flush is needed between iterations.

Every leg - the GPU arm, its `cpu_baseline`, its `gpu_aten_baseline`, and `--impl reference` - runs the SAME
tensors: rank r's shard is drawn on the CPU from `torch.Generator().manual_seed(1234 + r)` (S first, then T).  The
GPU arm checks its two loss values against the CPU leg's (1e-5) before it prints.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_PER_GPU, C, H, W = 16, 150, 128, 128
CGD = dict(group_size=10, alpha=3, tau=2)
METRIC = 'distill-loss fwd+bwd Mpixel/s'
UNIT = 'Mpixel/s'
WORKLOAD = ('cfg5 shard: CGDLoss(g=10,tau=2,alpha=3)+CDLoss fwd+bwd on logits %dx%dx%dx%d fp32 per GPU '
            '(global B=128 at 8 GPUs)' % (B_PER_GPU, C, H, W))
FALLBACK_HBM_GBS = 6650.0
# one dict for both arms (the driver compares them): what is computed, on which tensors
CONFIG = {'workload': WORKLOAD, 'per_gpu_shape': [B_PER_GPU, C, H, W], 'losses': ['CGDLoss', 'CDLoss'],
          'inputs': 'rank r: torch.Generator().manual_seed(1234 + r) on the CPU, S = randn, then T = randn, fp32',
          'l2': 'inputs (2 x 157 MB per GPU) exceed the 126 MB L2; no flush'}


def make_inputs(rank):
    """The shard of `rank`, drawn on the CPU so that every leg (GPU, CPU port, ATen-on-GPU) sees the same numbers."""
    import torch
    g = torch.Generator().manual_seed(1234 + rank)
    s = torch.randn(B_PER_GPU, C, H, W, generator=g)
    t = torch.randn(B_PER_GPU, C, H, W, generator=g)
    return s, t


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the CPU oracle leg (profiling runs)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-aten', action='store_true', help='skip the ATen-chain-on-GPU baseline leg')
    ap.add_argument('--extra', action='store_true', help='also time the other BASELINE configs (N=1 only)')
    ap.add_argument('--unfused', action='store_true', help='one launch per loss (no dispatcher batching)')
    ap.add_argument('--eager', action='store_true', help='time eager launches instead of replaying the captured step')
    return ap.parse_args()


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, window=None):
        """window = (t0, t1) in perf_counter seconds: only samples that arrived inside it count (the GPU was under
        the benchmark's load then)."""
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for stamp, ln in self.lines:
            if window is not None and not (window[0] <= stamp <= window[1]):
                continue
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': smax,
                'reasons': sorted(reasons), 'samples': len(sm), 'power_w_max': max(power) if power else None}


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, copy burst)'
    except Exception:
        return FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def profiled_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f).get('dominant_kernel_dram_bytes_per_launch')
    except Exception:
        return None


# ------------------------------------------------------------------ CPU oracle legs
def cpu_oracle_step_time(s, t, steps, warmup, threads=None):
    """Reference algorithm (oracle port of losses.py) on the host cores: CGD+CD fwd+bwd on the shard (s, t).
    Returns (seconds per timed step, (cgd loss, cd loss))."""
    import torch
    import oracle
    if threads:
        torch.set_num_threads(threads)
    gt = torch.zeros(s.shape[0], 1, H, W, dtype=torch.long)
    crits = [oracle.make_preset('CGDLoss', **CGD), oracle.make_preset('CDLoss')]
    times = []
    for i in range(warmup + steps):
        x = s.clone().requires_grad_(True)
        t0 = time.perf_counter()
        l_cgd, l_cd = crits[0](x, t, gt, 1), crits[1](x, t, gt, 1)
        (l_cgd + l_cd).backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, (float(l_cgd.detach()), float(l_cd.detach()))


def gpu_aten_step_time(crits_kwargs, s, t, gt, steps=5, warmup=2, loss_fn=None):
    """The reference's ATen op chain (oracle port of losses.py:95-113 + autograd) on the SAME GPU: what the drop-in
    replaces when the reference itself runs on a B200 ("the true kernel to beat").  CUDA events.
    crits_kwargs: [(preset name, ctor kwargs)]; or loss_fn(x) -> scalar for a chain that is not a preset."""
    import torch
    import oracle
    crits = [oracle.make_preset(name, **kw) for name, kw in (crits_kwargs or [])]
    x = s.detach().clone().requires_grad_(True)

    def step():
        x.grad = None
        total = loss_fn(x) if loss_fn is not None else None
        for c in crits:
            v = c(x, t, gt, 1)
            total = v if total is None else total + v
        total.backward()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def run_reference_arm(args, rank):
    """--impl reference: the reference's own algorithm on the host CPU (oracle port; the reference is a
    Python package that cannot travel to the GPU box), every host thread, bounded sample per step."""
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    s, t = make_inputs(0)                 # the whole shard of rank 0: the tensors the GPU arm runs
    times, losses = cpu_oracle_step_time(s, t, args.steps, args.warmup)
    total = sum(times)
    value = B_PER_GPU * H * W * len(times) / total / 1e6
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': dict(CONFIG),
        'timing': 'host perf_counter, CPU only, mean over the timed steps',
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': f'each step = CGD+CD fwd+bwd on the whole shard of rank 0 ({B_PER_GPU}x{C}x{H}x{W} fp32, '
                                   f'the GPU arm\'s tensors), oracle port of losses.py on torch CPU, every host thread; '
                                   f'the reference is Python on top of mmcv (absent): it cannot travel to the GPU box'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'loss_values': {'cgd': losses[0], 'cd': losses[1]},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
def _note(rank, msg):
    """progress marker on stderr (a multi-GPU run that stalls shows where)"""
    if os.environ.get('SD_BENCH_VERBOSE', '') or int(os.environ.get('WORLD_SIZE', '1')) > 1:
        print(f'[bench rank {rank} +{time.perf_counter() - _T0:6.1f}s] {msg}', file=sys.stderr, flush=True)


_T0 = time.perf_counter()


def _arm_watchdog(seconds):
    """A rank that stalls (a peer died, a collective never completes) must not hold the box: dump the Python stacks
    and exit non-zero after `seconds`."""
    import faulthandler
    faulthandler.dump_traceback_later(seconds, exit=True)


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import segdistill_b200 as sd
    from segdistill_b200 import _cabi
    from segdistill_b200 import dist as sdist

    _arm_watchdog(int(os.environ.get('SD_BENCH_WATCHDOG_S', '420')))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        _note(rank, 'init_process_group(nccl)')
        dist.init_process_group('nccl', device_id=dev)
        _note(rank, 'process group up')
    assert _cabi.load().sd_device_check() == 0, 'not an sm_100 device'

    hS, hT = make_inputs(rank)                      # CPU, seeded: every leg runs these tensors
    hS, hT = hS.pin_memory(), hT.pin_memory()
    S = hS.to(dev).requires_grad_(True)
    T = hT.to(dev)
    gt = torch.zeros(B_PER_GPU, 1, H, W, dtype=torch.long, device=dev)
    dl = sd.DistillationLoss([
        {'student_layer': 'decode_head.linear_pred', 'teacher_layer': 'decode_head.linear_pred',
         'loss_name': 'CGDLoss', 'loss_config': dict(CGD)},
        {'student_layer': 'decode_head', 'teacher_layer': 'decode_head', 'loss_name': 'CDLoss', 'loss_config': {}}])
    dl.batch_pairs = not args.unfused
    numel = S.numel()
    # N > 1: the path's only collective is the all-reduce of the loss scalars for logging.  The reference issues one
    # per log variable per step (SD_structure.py:137-142); here every step appends its scalars to a device-resident
    # ring (no launch of its own: the append rides on the backward's scaling launch, sd_scale_grad_log; captured with
    # the step) and the ring is all-reduced ONCE per LOG_INTERVAL steps -
    # the interval at which the reference's logger reads them (default_runtime.py:2-7).  No NCCL kernel sits between
    # two steps' loss kernels; the flush of the steps timed here is inside the timed region.
    LOG_INTERVAL = 50
    logs = sdist.DeferredLogs(['loss_cgd', 'loss_cd'], interval=LOG_INTERVAL, device=dev) if world > 1 else None
    log_stream = torch.cuda.Stream() if world > 1 else None

    def feats(x):
        return {'decode_head.linear_pred': x, 'decode_head': x}
    fS, fT = feats(S), feats(T)

    def compute(record=None):
        S.grad = None
        if record:
            record[0].record()
        out = dl(fS, fT, gt, 1, None, None)
        if record:
            record[1].record()
            _cabi.last_kernel_of_step = _cabi.last_kernel()
        l1, l2 = out.values()
        if logs is not None:
            logs.push([l1, l2], in_backward=True)      # device-side append, no collective, no launch of its own: it rides
                                                       # on the backward's scaling launch (sd_scale_grad_log)
        (l1 + l2).backward()
        if logs is not None:
            logs.join()
        if record:
            record[2].record()
        return l1, l2

    flushed, in_flight = [], []

    def flush_logs():
        """one all-reduce + one (asynchronous) D2H for all the steps since the last flush; the launch stream is ordered
        behind the collective, the host reads the values later (collect_logs)"""
        if logs is not None:
            in_flight.append(logs.flush_start())

    def collect_logs():
        while in_flight:
            flushed.extend(logs.flush_finish(in_flight.pop(0)))

    def step(record=None):
        return compute(record)

    def sync_all():
        torch.cuda.synchronize()
        collect_logs()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from here to the end of the timed region (nvidia-smi needs ~0.2 s to come up:
    # started right before a 3 ms timed loop it would sample nothing and contend for the driver inside it)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    _note(rank, 'warm-up')
    for _ in range(args.warmup):
        step()
    sync_all()
    _note(rank, 'warm-up done')

    # The step (dispatcher forward + autograd backward [+ the scalar all-reduce]) is a fixed launch sequence:
    # capture it once into a CUDA graph and replay it.  Eager per-step numbers are measured beside it.
    graph, graph_note = None, 'eager launches'
    if not args.eager:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            sync_all()
            # capture on the stream the warm-up ran on: its zero-filled workspace exists already (a fresh
            # capture stream would put the one-time workspace allocation + fill into every replay).  The scalar
            # append of the step's loss scalars to the log ring (N > 1) is part of the captured step.
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_, stream=side):
                static_losses = compute()
            graph, graph_note = [g_], 'CUDA graph replay of the captured module-API step'
            for i in range(4):
                graph[0].replay()
            flush_logs()
            sync_all()
            _note(rank, 'graph captured')
        except Exception as exc:          # capture is an optimisation, never a requirement
            graph, graph_note = None, f'eager launches (graph capture failed: {type(exc).__name__})'
            torch.cuda.synchronize()
    _note(rank, 'timed region')
    t_load0 = time.perf_counter()
    # ---- timed region: exactly K steps, device time
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _cabi.launch_count()
    sync_all()
    # eager pass: per-step events around the loss kernels (roofline) and the host-driven step time
    t_begin.record()
    for i in range(args.steps):
        l1, l2 = step(evs[i])
        if logs is not None and (i + 1) % LOG_INTERVAL == 0:
            flush_logs()
    flush_logs()                             # (N > 1) the scalars of the timed steps are reduced inside the timed region
    t_end.record()
    sync_all()
    launches = _cabi.launch_count() - launches0
    eager_ms = t_begin.elapsed_time(t_end) / args.steps
    # the eager figure proper: the same step without the per-step event records, over enough steps (>= 200) that one
    # nvidia-smi poll of the clock sampler (every 100 ms, it holds the driver for milliseconds) cannot dominate a
    # 2-8 ms window
    n_eager = max(args.steps, 200)
    t_begin.record()
    for i in range(n_eager):
        step()
        if logs is not None and (i + 1) % LOG_INTERVAL == 0:
            flush_logs()
    flush_logs()
    t_end.record()
    sync_all()
    eager_ms_events = eager_ms
    eager_ms = t_begin.elapsed_time(t_end) / n_eager
    if world > 1:
        tt = torch.tensor([eager_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        eager_ms = tt.item()
    if graph is not None:
        # timed region of the headline number: exactly K replays of the captured step
        per_step_launches = launches // args.steps
        t_begin.record()
        for i in range(args.steps):
            graph[0].replay()
            if logs is not None and (i + 1) % LOG_INTERVAL == 0:
                flush_logs()
        t_flush = torch.cuda.Event(enable_timing=True)
        t_flush.record()
        flush_logs()                         # the timed region ends when the steps' scalars have been reduced and read
        t_end.record()
        sync_all()
        final_flush_ms = t_flush.elapsed_time(t_end) if logs is not None else None
        l1, l2 = static_losses
        launches = per_step_launches * args.steps
    elapsed_ms = t_begin.elapsed_time(t_end)
    # the timed region lasts a few milliseconds, nvidia-smi samples every 100 ms: keep the SAME load running (untimed)
    # for ~0.6 s so that the clock / throttle samples are taken under it
    t_hold = time.perf_counter() + 0.6
    while time.perf_counter() < t_hold:
        for _ in range(50):
            if graph is not None:
                graph[0].replay()
            else:
                compute()
        torch.cuda.synchronize()
    clocks = sampler.stop(window=(t_load0, time.perf_counter())) if rank == 0 else None
    per_rank = None
    if world > 1:
        # every rank's own device time of the region (and of its K steps without the closing flush): the spread between
        # the GPUs of the box is part of the max-over-ranks figure
        steps_only = t_begin.elapsed_time(t_flush) if graph is not None else elapsed_ms
        mine = torch.tensor([elapsed_ms, steps_only], device=dev)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank = {'region_ms_per_step': [round(t[0].item() / args.steps, 6) for t in every],
                    'steps_only_ms_per_step': [round(t[1].item() / args.steps, 6) for t in every]}
        tt = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = tt.item()
    assert _cabi.workspace_error_flag() == 0
    fwd_ms = statistics.median(e[0].elapsed_time(e[1]) for e in evs)   # the loss kernel(s) of one eager step (+ host gaps);
    # median over the K steps: one poll of the clock sampler inside a step's window is milliseconds
    # the dominant kernel alone: the same fused launch through the C ABI, back to back on this stream
    kern_ms = None
    if dl.batch_pairs:
        with torch.no_grad():
            for _ in range(3):
                _cabi.kl_rows_multi(S, T, (CGD['group_size'], 1), (float(CGD['tau']), 1.0), (float(CGD['alpha']), 1.0))
            ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ka.record()
            for _ in range(args.steps):
                _cabi.kl_rows_multi(S, T, (CGD['group_size'], 1), (float(CGD['tau']), 1.0), (float(CGD['alpha']), 1.0))
            kb.record()
            torch.cuda.synchronize()
            kern_ms = ka.elapsed_time(kb) / args.steps
    bwd_ms = statistics.median(e[1].elapsed_time(e[2]) for e in evs)
    ms_per_step = elapsed_ms / args.steps
    value = world * B_PER_GPU * H * W / (ms_per_step * 1e-3) / 1e6

    _note(rank, f'timed region done: {elapsed_ms / args.steps:.4f} ms per step')
    # ---- e2e: pinned host buffers -> H2D -> modules -> D2H of the loss scalars, every step.  The inputs are double-buffered:
    #      the H2D copies of step k + 1 run on a copy stream while step k's kernels run and its result is read (what a
    #      training loop with a prefetching loader does); every step's copies and its result read are inside the timed
    #      region, n copies for n steps.
    e2e = None
    if not args.no_e2e:
        bufs = [(torch.empty_like(S).requires_grad_(True), torch.empty_like(T)) for _ in range(2)]
        copy_stream = torch.cuda.Stream()
        ready = [torch.cuda.Event() for _ in range(2)]      # the buffer pair holds the step's inputs
        free = [torch.cuda.Event() for _ in range(2)]       # the step that used the buffer pair is through with it

        def start_copy(i):
            k = i & 1
            copy_stream.wait_event(free[k])                 # (never recorded yet: returns at once)
            with torch.cuda.stream(copy_stream), torch.no_grad():
                bufs[k][0].copy_(hS, non_blocking=True)
                bufs[k][1].copy_(hT, non_blocking=True)
                ready[k].record(copy_stream)

        def e2e_step(i, prefetch):
            k = i & 1
            dS_in, dT_in = bufs[k]
            dS_in.grad = None
            torch.cuda.current_stream().wait_event(ready[k])
            if prefetch:
                start_copy(i + 1)
            losses = dl(feats(dS_in), feats(dT_in), gt, 1, None, None)
            total, log_vars = sdist.parse_losses(losses)     # one packed all-reduce + ONE D2H read per step
            total.backward()
            free[k].record()
            return log_vars

        def e2e_run(n):
            start_copy(0)
            for i in range(n):
                e2e_step(i, i + 1 < n)

        n_e2e = max(4, min(args.steps, 10))
        e2e_run(2)
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e2e_run(n_e2e)
        b.record()
        sync_all()
        ems = a.elapsed_time(b)
        if world > 1:
            tt = torch.tensor([ems], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ems = tt.item()
        e2e = {'value': world * B_PER_GPU * H * W / (ems / n_e2e * 1e-3) / 1e6, 'unit': UNIT,
               'h2d_bytes_per_step': 2 * numel * 4, 'd2h_bytes_per_step': 3 * 4, 'steps': n_e2e,
               'ms_per_step': ems / n_e2e,
               'pipelining': 'inputs double-buffered: the H2D of step k+1 (copy stream) overlaps the kernels and the result '
                             'read of step k; every step copies its own inputs and reads its own result inside the timed region'}
        del bufs

    extra = None
    if args.extra and world == 1:
        extra = extra_configs(dev)

    # ---- CPU baseline (rank 0, N=1 only): oracle port on a bounded sample
    cpu, loss_check, aten = None, None, None
    gpu_losses = {'cgd': float(l1.detach()), 'cd': float(l2.detach())}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        times, cpu_losses = cpu_oracle_step_time(hS, hT, steps=4, warmup=1, threads=cores)
        mean = sum(times) / len(times)
        cpu = {'value': B_PER_GPU * H * W / mean / 1e6, 'unit': UNIT, 'cores': torch.get_num_threads(),
               'kind': 'port',
               'sample': f'CGD+CD fwd+bwd on the whole shard ({B_PER_GPU}x{C}x{H}x{W} fp32, the tensors the GPU arm ran), '
                         f'mean of 4 steps after 1 warm-up ({sum(times):.1f} s of CPU work), oracle port of losses.py on '
                         f'torch CPU, every host thread'}
        # the benchmarked step computed the reference's numbers (north_star: loss rel. err <= 1e-5)
        errs = {'cgd': abs(gpu_losses['cgd'] - cpu_losses[0]) / abs(cpu_losses[0]),
                'cd': abs(gpu_losses['cd'] - cpu_losses[1]) / abs(cpu_losses[1])}
        loss_check = {'cpu_port': {'cgd': cpu_losses[0], 'cd': cpu_losses[1]}, 'rel_err': errs, 'tol': 1e-5,
                      'ok': max(errs.values()) <= 1e-5}
        assert loss_check['ok'], f'benchmarked losses differ from the CPU port: {loss_check}'
    if rank == 0 and world == 1 and not args.no_aten:
        ms = gpu_aten_step_time([('CGDLoss', dict(CGD)), ('CDLoss', {})], S, T, gt)
        aten = {'ms_per_step': ms, 'value': B_PER_GPU * H * W / (ms * 1e-3) / 1e6, 'unit': UNIT,
                'what': 'the reference\'s op chain (oracle port of losses.py:95-113: div, log_softmax, softmax, kl_div, autograd '
                        'backward) as ATen CUDA kernels on this GPU, same tensors, CUDA events, 5 steps after 2 warm-ups'}
        torch.cuda.empty_cache()

    if rank == 0:
        peak, peak_src = measured_peak()
        cd_bytes = 12.0 * numel                      # read S + read T + write dS, fp32
        launches_per_step = 1 if dl.batch_pairs else 2
        achieved = (cd_bytes / (kern_ms * 1e-3) / 1e9) if kern_ms else launches_per_step * cd_bytes / (fwd_ms * 1e-3) / 1e9
        kname = (_cabi.last_kernel_of_step + ' (CGD+CD fused, one launch per step)' if dl.batch_pairs
                 else 'kl_rows_* (mean of the CGD and CD launches)')
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(CONFIG),
            'timing': 'CUDA events on the launch stream, max over ranks',
            'launch': graph_note,
            'parallelism': (f'batch-sharded x{world}; loss scalars appended to a device ring every step, one all-reduce per '
                            f'{LOG_INTERVAL} steps (and at the end of the timed region)' if world > 1 else 'single GPU'),
            # the same step driven eagerly through the module API (what mmcv's runner does), device-timed
            'eager': {'ms_per_step': eager_ms, 'value': world * B_PER_GPU * H * W / (eager_ms * 1e-3) / 1e6, 'unit': UNIT,
                      'steps': n_eager, 'ms_per_step_with_event_records': eager_ms_events},
            'gpu_aten_baseline': aten, 'loss_check': loss_check,
            'melem_per_s': world * numel / (ms_per_step * 1e-3) / 1e6,
            'hbm_gbs_step': world * launches_per_step * cd_bytes / (ms_per_step * 1e-3) / 1e9,
            'kernel_ms': {'dominant_kernel': kern_ms, 'loss_kernels_fwd_eager': fwd_ms, 'backward_eager': bwd_ms,
                          'eager_ms_per_step': eager_ms},
            'roofline': {'bound': 'hbm', 'kernel': kname,
                         'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'peak_source': peak_src, 'frac_of_8TBs_nominal': achieved / 8000.0,
                         'algorithmic_bytes_per_launch': cd_bytes, 'traffic': profiled_traffic(),
                         'timing': 'CUDA events around back-to-back launches of the kernel on the launch stream'},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'cpu_baseline': cpu,
            'loss_values': gpu_losses,
        }
        if flushed:
            line['reduced_log_sample'] = flushed[-1]
        if graph is not None and final_flush_ms is not None:
            # rank 0's device time of the flush that closes the timed region (all-reduce of the ring + its read-back,
            # the wait for the slowest rank included): the part of ms_per_step x steps that is not the steps themselves
            line['log_flush'] = {'final_flush_ms': final_flush_ms, 'per_step_ms_at_this_K': final_flush_ms / args.steps,
                                 'interval_steps': LOG_INTERVAL}
        if per_rank is not None:
            line['per_rank'] = per_rank
        if extra:
            line['extra'] = extra
        print(json.dumps(line), flush=True)
    _note(rank, 'done')
    if world > 1:
        dist.destroy_process_group()


def extra_configs(dev):
    """Other BASELINE configs (parity-test cases, timed for information only)."""
    import torch
    import segdistill_b200 as sd
    from segdistill_b200 import _cabi
    out = {}

    def timeit(fn, n=20):
        """(eager ms, CUDA-graph replay ms) per call of fn (forward + backward through the module API)."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        eager = a.elapsed_time(b) / n
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            a.record()
            for _ in range(n):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            return eager, a.elapsed_time(b) / n
        except Exception:
            torch.cuda.synchronize()
            return eager, eager

    def fwd_bwd(crit, s, t):
        def f():
            s.grad = None
            crit(s, t, None, 1).backward()
        return f

    def pair(shape, dtype):
        s = torch.randn(shape, device=dev).to(dtype).requires_grad_(True)
        t = torch.randn(shape, device=dev).to(dtype)
        return s, t

    import oracle
    peak, _ = measured_peak()

    def aten_ms(presets, s, t, gt_hw=None, loss_fn=None):
        """the reference's ATen chain for the same loss on this GPU (ms per fwd+bwd), or None when it does not fit"""
        hw = gt_hw or tuple(s.shape[2:])
        gt = torch.zeros(s.shape[0], 1, *hw, dtype=torch.long, device=dev)
        try:
            return gpu_aten_step_time(presets, s, t, gt, steps=3, warmup=1, loss_fn=loss_fn)
        except torch.OutOfMemoryError:
            return None
        finally:
            torch.cuda.empty_cache()

    for name, crit, shape, dtype, presets in (
            ('cfg1_cd_2x150x64x64_f32', sd.CDLoss(), (2, 150, 64, 64), torch.float32, [('CDLoss', {})]),
            ('cfg3_cd_16x150x128x128_bf16', sd.CDLoss(), (16, 150, 128, 128), torch.bfloat16, [('CDLoss', {})]),
            ('cfg3_pd_16x150x128x128_bf16', sd.PDLoss(), (16, 150, 128, 128), torch.bfloat16, [('PDLoss', {})]),
            ('cfg3_pd_16x150x128x128_f32', sd.PDLoss(), (16, 150, 128, 128), torch.float32, [('PDLoss', {})]),
            ('cfg4_cd+mse_fused_16x512x64x64_f32', sd.CDMSELoss(alpha=1, tau=4), (16, 512, 64, 64), torch.float32,
             [('CGDLoss', dict(group_size=1, alpha=1, tau=4))]),
            ('cfg4_mse_16x512x64x64_f32', sd.FeatureMSELoss(), (16, 512, 64, 64), torch.float32, [])):
        s, t = pair(shape, dtype)
        eager, ms = timeit(fwd_bwd(crit, s, t))
        nbytes = 3 * s.numel() * s.element_size()
        mse_fn = (lambda x, t=t: oracle.mse_loss_torch(x, t, 1.0)) if 'mse' in name else None
        out[name] = {'ms': ms, 'ms_eager': eager, 'mpixel_s': shape[0] * shape[2] * shape[3] / ms / 1e3,
                     'gbs': nbytes / ms / 1e6, 'frac_of_measured_peak': nbytes / ms / 1e6 / peak,
                     'aten_chain_ms': aten_ms(presets, s, t, loss_fn=mse_fn)}
    stages = [(16, 32, 128, 128), (16, 64, 64, 64), (16, 160, 32, 32), (16, 256, 16, 16)]
    pairs = [pair(sh, torch.float32) for sh in stages]
    crit = sd.CGDLoss()

    def cfg2():
        for s, t in pairs:
            s.grad = None
            crit(s, t, None, 1).backward()
    eager, ms = timeit(cfg2)
    nbytes = sum(3 * s.numel() * 4 for s, _ in pairs)
    # the same through the dispatcher: four `distillation` entries on four layers -> ONE grouped launch per step (f3)
    dl2 = sd.DistillationLoss([{'student_layer': f's{k}', 'teacher_layer': f's{k}', 'loss_name': 'CGDLoss', 'loss_config': {}}
                               for k in range(4)])
    fs2 = {f's{k}': s for k, (s, _) in enumerate(pairs)}
    ft2 = {f's{k}': t for k, (_, t) in enumerate(pairs)}

    kname = {}

    def cfg2_grouped():
        for s, _ in pairs:
            s.grad = None
        losses2 = dl2(fs2, ft2, None, 1, None, None)
        kname['fwd'] = _cabi.last_kernel()
        sum(losses2.values()).backward()
    eager_g, ms_g = timeit(cfg2_grouped)
    out['cfg2_cgd_4stages_b16_f32_grouped'] = {'ms': ms_g, 'ms_eager': eager_g, 'gbs': nbytes / ms_g / 1e6,
                                               'frac_of_measured_peak': nbytes / ms_g / 1e6 / peak,
                                               'launches_per_step': 1, 'kernel': kname.get('fwd')}
    aten2 = [aten_ms([('CGDLoss', {})], s, t) for s, t in pairs]
    out['cfg2_cgd_4stages_b16_f32'] = {'ms': ms, 'ms_eager': eager, 'mpixel_s': sum(s.shape[0] * s.shape[2] * s.shape[3] for s, _ in pairs) / ms / 1e3,
                                       'gbs': nbytes / ms / 1e6, 'frac_of_measured_peak': nbytes / ms / 1e6 / peak,
                                       'aten_chain_ms': sum(aten2) if all(v is not None for v in aten2) else None}
    # the reference's real training path: logits at 1/4 resolution resized to the 512x512 labels;
    # fused = bilinear up-sampling inside the loss kernels, host = F.interpolate first (what the reference does)
    for name, cls, shape in (('f1_cgd_resize4_2x150x128x128_f32', sd.CGDLoss, (2, 150, 128, 128)),
                             ('f1_pd_resize4_2x150x128x128_f32', sd.PDLoss, (2, 150, 128, 128)),
                             ('f1_cd_resize4_16x150x128x128_f32', sd.CDLoss, (16, 150, 128, 128))):
        s, t = pair(shape, torch.float32)
        gt = torch.zeros(shape[0], 1, 4 * shape[2], 4 * shape[3], dtype=torch.long, device=dev)
        rec = {}
        for tag, fuse in (('fused', True), ('host_resize', False)):
            crit = cls()
            crit.fuse_resize = fuse

            def f(crit=crit):
                s.grad = None
                crit(s, t, gt, 1).backward()
            eager, ms = timeit(f, n=10)
            rec[tag + '_ms'] = ms
        hi = shape[0] * shape[1] * 16 * shape[2] * shape[3]
        rec['mpixel_s'] = shape[0] * 16 * shape[2] * shape[3] / rec['fused_ms'] / 1e3
        rec['upsampled_gelem_s'] = hi / rec['fused_ms'] / 1e6
        rec['speedup_vs_host_resize'] = rec['host_resize_ms'] / rec['fused_ms']
        # these kernels regenerate every up-sampled value twice (statistics, gradient) and take its exponential for S and
        # for T each time: 4 ex2 per up-sampled value.  The bound is the MUFU pipe (16 ex2 / clk / SM), not HBM.
        mufu_peak = 148 * 16 * 1.965e9
        rec['roofline'] = {'bound': 'mufu', 'ex2_per_upsampled_value': 4, 'achieved_gex2_s': 4 * hi / rec['fused_ms'] / 1e6,
                           'peak_gex2_s': mufu_peak / 1e9, 'frac': 4 * hi / (rec['fused_ms'] * 1e-3) / mufu_peak,
                           'peak_source': '148 SMs x 16 MUFU.EX2 / clk x 1.965 GHz (B300_MICROARCH.md; sm clock from MEASURED_PEAKS.json)'}
        if shape[0] <= 2:                      # the reference chain on the resized maps (16x the elements): small batch only
            rec['aten_chain_ms'] = aten_ms([(cls.__name__, {})], s, t, gt_hw=(4 * shape[2], 4 * shape[3]))
        out[name] = rec
    # IFVDLoss on the training shape: logits 2x150x128x128, labels 512x512 in 32-pixel blocks with an ignore band
    s, t = pair((2, 150, 128, 128), torch.float32)
    lab = torch.randint(0, 150, (2, 1, 16, 16), device=dev).repeat_interleave(32, 2).repeat_interleave(32, 3)
    lab[:, :, 100:140, :200] = 255
    ifvd = sd.IFVDLoss()

    def f_ifvd():
        s.grad = None
        ifvd(s, t, lab, 1).backward()
    eager, ms = timeit(f_ifvd)
    out['f2_ifvd_2x150x128x128_f32'] = {'ms': ms, 'ms_eager': eager, 'mpixel_s': 2 * 128 * 128 / ms / 1e3}
    # the correlation extension (tensor cores): algorithmic traffic 16 B/element fp32 (S, T read; S read again; dS written)
    for name, g, shape, dtype in (('corr_g10_16x150x128x128_bf16', 10, (16, 150, 128, 128), torch.bfloat16),
                                  ('corr_g10_16x150x128x128_f32', 10, (16, 150, 128, 128), torch.float32),
                                  ('corr_g256_16x512x64x64_bf16', 256, (16, 512, 64, 64), torch.bfloat16)):
        s, t = pair(shape, dtype)
        eager, ms = timeit(fwd_bwd(sd.CGDCorrLoss(group_size=g), s, t))
        nbytes = 4 * s.numel() * s.element_size()
        gpb = -(-shape[1] // g)
        flops = 6.0 * shape[0] * gpb * g * g * shape[2] * shape[3]        # useful: 2 Grams + the gradient GEMM
        out[name] = {'ms': ms, 'ms_eager': eager, 'gbs': nbytes / ms / 1e6, 'frac_of_measured_peak': nbytes / ms / 1e6 / peak,
                     'useful_tflops': flops / ms / 1e9}
    return out


def main():
    args = parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.gpus > 1 and world == 1 and args.impl == 'ours':
        # launched without torchrun: re-exec under it (one rank per GPU, NCCL)
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', str(29500 + os.getpid() % 1000),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == 'reference':
        run_reference_arm(args, rank)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
