"""world_size-2 gloo tests of the data-parallel plumbing: batch sharding + one gradient all-reduce
gives the same update as the single-process step on the whole batch.  The per-rank model here is a small CPU stand-in
for the stem (the CUDA op has no CPU path); the plumbing under test is device-agnostic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Conv1d(4, 6, 3, padding=1), torch.nn.GELU(), torch.nn.Conv1d(6, 2, 3, stride=2, padding=1))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qasr_ijcnlp_b200 import dp

        torch.manual_seed(100 + rank)  # deliberately different init per rank: broadcast must fix it
        model = torch.nn.Sequential(torch.nn.Conv1d(4, 6, 3, padding=1), torch.nn.GELU(),
                                    torch.nn.Conv1d(6, 2, 3, stride=2, padding=1))
        if rank == 0:
            model.load_state_dict(_model().state_dict())
        dp.broadcast_parameters(model, src=0)
        g = torch.Generator().manual_seed(9)
        x = torch.randn(6, 4, 16, generator=g)  # the global batch, identical on every rank
        lo, hi = dp.shard_range(x.shape[0], rank, world)
        bucket = dp.GradBucket(model.parameters())
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        # loss is a SUM over utterances divided by the global batch -> mean over ranks of (world * local mean-share)
        loss = model(x[lo:hi]).square().sum() / x.shape[0] * world
        loss.backward()
        pending = bucket.allreduce_mean(async_op=True)
        if pending is not None:
            pending.wait()
        opt.step()
        q.put((rank, (lo, hi), [p.detach().numpy().copy() for p in model.parameters()]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_step_equals_single_process_step():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=150) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [(0, 3), (3, 6)]
    # single-process reference step on the whole batch
    model = _model()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(6, 4, 16, generator=g)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    (model(x).square().sum() / x.shape[0]).backward()
    opt.step()
    for rank_params in (res[0][2], res[1][2]):
        for a, b in zip(rank_params, model.parameters()):
            assert abs(a - b.detach().numpy()).max() <= 1e-6
    for a, b in zip(res[0][2], res[1][2]):
        assert (a == b).all()  # ranks stay bit-identical after the all-reduce
