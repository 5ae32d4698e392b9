"""CPU, world_size 2, gloo: the host-side sharding logic of the multi-GPU rollout —
contiguous batch shards, one all-gather of per-sample losses, global LpLoss mean."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fourierflow_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, n_steps, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = D.init_from_env(backend="gloo")
    full = torch.arange(n_steps * n_global, dtype=torch.float32).reshape(n_steps, n_global) * 0.01 + 0.5
    lo, hi = D.shard_bounds(n_global, r, w)
    local = full[:, lo:hi].contiguous()
    loss, step = D.rollout_loss(local, n_global)
    gathered = D.gather_sample_losses(local, n_global)
    # the side-stream gatherer bench.py uses (preallocated buffers; here on CPU: same collective, same layout)
    g = D.LossGatherer(n_steps, hi - lo, n_global, torch.device("cpu"))
    for _ in range(2):                       # reused buffers: a second submit must give the same answer
        g.submit(local)
        again = g.result()
    assert torch.equal(again, gathered), (again, gathered)
    q.put((rank, loss.item(), step.tolist(), gathered.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def _run(n_global, n_steps=3, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_global, n_steps, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    full = torch.arange(n_steps * n_global, dtype=torch.float32).reshape(n_steps, n_global) * 0.01 + 0.5
    for rank, loss, step, gathered in res:
        assert torch.allclose(torch.tensor(gathered), full)
        assert torch.allclose(torch.tensor(step), full.mean(dim=1))
        assert abs(loss - full.mean(dim=1).sum().item()) < 1e-5


def test_even_shards():
    _run(8)


def test_ragged_shards():
    _run(5)


def test_shard_bounds_cover_batch():
    for n in (0, 1, 7, 32, 256):
        for w in (1, 2, 3, 8):
            cuts = [D.shard_bounds(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1


def _grad_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    D.init_from_env(backend="gloo")
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2))]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(5, dtype=torch.float32) * (rank + 1)
    # params[2] has no gradient on any rank: skipped
    n = D.allreduce_gradients(params)
    q.put((rank, n, params[0].grad.clone(), params[1].grad.clone(), params[2].grad))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_averages_over_ranks():
    """Data-parallel training (the reference's dormant DDP, commands/train.py:83-84): one flat all-reduce, mean over ranks."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, n, g0, g1, g2 in res:
        assert n == 17 and g2 is None
        assert torch.allclose(g0, torch.full((3, 4), 1.5))
        assert torch.allclose(g1, torch.arange(5, dtype=torch.float32) * 1.5)
