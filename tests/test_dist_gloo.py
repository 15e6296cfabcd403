"""N>1 path on CPU: world_size-2 gloo.  The batch shards across ranks, each rank computes its
samples' loss locally (oracle stands in for the kernel on CPU) and the only collective is the packed
scalar all-reduce; mean over ranks == single-process global-batch loss."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import oracle
    from segdistill_b200 import dist as sdist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        s = torch.randn(8, 20, 8, 8, generator=g)
        t = torch.randn(8, 20, 8, 8, generator=g)
        gt = torch.zeros(8, 1, 8, 8, dtype=torch.long)
        lo, hi = sdist.shard_bounds(8, rank, world)
        crits = {'loss_cgd': oracle.make_preset('CGDLoss'), 'loss_cd': oracle.make_preset('CDLoss')}
        local = {k: c(s[lo:hi], t[lo:hi], gt[lo:hi], 1) for k, c in crits.items()}
        local['acc_dummy'] = torch.tensor(float(rank))
        total, logs = sdist.parse_losses(local)
        # the deferred form: three steps appended locally, one all-reduce for all of them at the flush
        dl = sdist.DeferredLogs(['loss_cgd', 'loss_cd', 'acc_dummy'], interval=4)
        for step in range(3):
            dl.push([local['loss_cgd'] * (step + 1), local['loss_cd'], local['acc_dummy']])
        deferred = dl.flush()
        assert dl.flush() == []
        for step in range(6):                  # more steps than slots: the newest `interval` survive
            dl.push([local['loss_cgd'] * (step + 1), local['loss_cd'], local['acc_dummy']])
        wrapped = dl.flush()
        glob = {k: float(oracle.make_preset(k2)(s, t, gt, 1)) for k, k2 in (('loss_cgd', 'CGDLoss'), ('loss_cd', 'CDLoss'))}
        q.put((rank, float(total), dict(logs), glob, [dict(d) for d in deferred], [dict(d) for d in wrapped]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_packed_allreduce_matches_global_batch():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    res.sort()
    logs0, logs1 = res[0][2], res[1][2]
    assert logs0 == logs1                                   # every rank logs the same reduced values
    glob = res[0][3]
    for k in ('loss_cgd', 'loss_cd'):
        assert logs0[k] == pytest.approx(glob[k], rel=2e-6)  # mean of equal shards == global batch
    assert logs0['acc_dummy'] == pytest.approx(0.5)
    assert logs0['loss'] == pytest.approx(glob['loss_cgd'] + glob['loss_cd'], rel=2e-6)
    assert res[0][1] != res[1][1]                           # the differentiable totals stay local
    # DeferredLogs: the same rank-averaged values, one collective for the whole interval
    d0, d1 = res[0][4], res[1][4]
    assert d0 == d1 and len(d0) == 3
    for step, rec in enumerate(d0):
        assert rec['loss_cgd'] == pytest.approx((step + 1) * glob['loss_cgd'], rel=2e-6)
        assert rec['loss_cd'] == pytest.approx(glob['loss_cd'], rel=2e-6)
        assert rec['acc_dummy'] == pytest.approx(0.5)
        assert rec['loss'] == pytest.approx(rec['loss_cgd'] + rec['loss_cd'], rel=1e-6)
    w0 = res[0][5]
    assert len(w0) == 4 and [round(r['loss_cgd'] / glob['loss_cgd']) for r in w0] == [3, 4, 5, 6]
