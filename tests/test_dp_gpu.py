"""GPU tests of the data-parallel gradient collective (qw_grads_allreduce_p2p).  The single-GPU case exercises the
kernel's epoch / double-buffer logic; the 2-GPU case (skipped on a 1-GPU box) spawns one process per GPU and checks the
NVLink peer-memory all-reduce against NCCL, eagerly and from a replayed CUDA graph."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_world1_is_identity_and_graph_capturable(cuda):
    from qasr_ijcnlp_b200 import dp
    ar = dp.P2PGradAllReduce(9443, cuda)  # ragged: not a multiple of 4, several chunks
    g = torch.randn(9443, device=cuda)
    want = g.clone()
    for _ in range(3):  # epochs 1..3, both slots
        ar(g)
    assert torch.equal(g, want) and ar.status() == 0
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.graph(graph, stream=s):
        ar(g)
    for _ in range(4):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(g, want) and ar.status() == 0
    assert ar.epoch() == 3 + 4  # epoch advanced once per launch: 3 eager + 4 replays (capture itself does not execute)
    with pytest.raises(ValueError):
        ar(torch.zeros(5, device=cuda))
    big = dp.P2PGradAllReduce(1 << 20, cuda)  # 64 chunks
    h = torch.randn(1 << 20, device=cuda)
    want = h.clone()
    big(h); big(h)
    assert torch.equal(h, want) and big.status() == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from qasr_ijcnlp_b200 import dp
        n = 9440 + 35 * 385 + 3
        ar = dp.P2PGradAllReduce(n, dev)
        ok = True
        for it in range(5):
            g = torch.Generator(device=dev).manual_seed(100 * it + rank)
            x = torch.randn(n, device=dev, generator=g)
            ref = x.clone()
            dist.all_reduce(ref, op=dist.ReduceOp.SUM)
            ref /= world
            ar(x)
            ok = ok and (x - ref).abs().max().item() <= 1e-6
        # ranks must end bit-identical (fixed summation order)
        gathered = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(gathered, x)
        ok = ok and all(torch.equal(gathered[0], t) for t in gathered)
        # graph replay
        y = torch.full((n,), float(rank + 1), device=dev)
        graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.graph(graph, stream=s):
            ar(y)
        torch.cuda.synchronize()
        dist.barrier()
        want = float(sum(range(1, world + 1))) / world
        for _ in range(3):
            y.fill_(float(rank + 1))
            torch.cuda.synchronize()
            dist.barrier()
            graph.replay()
            torch.cuda.synchronize()
            ok = ok and bool((y - want).abs().max().item() <= 1e-6)
        ok = ok and ar.status() == 0
        # late rank (> 1 s): the one-shot kernel waits for it
        z = torch.full((n,), float(rank + 1), device=dev)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            import time
            time.sleep(1.6)
        ar(z)
        torch.cuda.synchronize()
        ok = ok and bool((z - want).abs().max().item() <= 1e-6)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_p2p_allreduce_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}


def test_fused_backward_world1_equals_plain_backward(cuda):
    """qw_conv1d_backward_dp with world == 1 is the plain backward (and the module hook is a no-op)."""
    from qasr_ijcnlp_b200 import QuantumConv1d
    torch.manual_seed(5)
    m = QuantumConv1d(80, 384, 3, padding=1, n_qubits=4).to(cuda)
    x = torch.randn(2, 80, 256, device=cuda)
    gy = torch.randn(2, 384, 256, device=cuda)
    ref = torch.autograd.grad(m(x), list(m.parameters()), gy)
    m.fuse_grad_allreduce()
    got = torch.autograd.grad(m(x), list(m.parameters()), gy)
    assert all(torch.equal(a, b) for a, b in zip(ref, got))


def _worker_fused(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from qasr_ijcnlp_b200 import QuantumConv1d
        ok = True
        for (C, S, B, L) in ((80, 1, 4, 3000), (384, 2, 2, 3000), (8, 1, 1, 64)):
            torch.manual_seed(11)  # same parameters on every rank
            plain = QuantumConv1d(C, 384, 3, stride=S, padding=1, n_qubits=4).to(dev)
            torch.manual_seed(11)
            fused = QuantumConv1d(C, 384, 3, stride=S, padding=1, n_qubits=4).to(dev).fuse_grad_allreduce()
            for it in range(3):  # both epoch parities
                g = torch.Generator(device=dev).manual_seed(1000 * it + rank)  # different data per rank
                x = torch.randn(B, C, L, device=dev, generator=g, requires_grad=True)
                gy = torch.randn(B, 384, (L + 2 - 3) // S + 1, device=dev, generator=g)
                ref = list(torch.autograd.grad(plain(x), [x] + list(plain.parameters()), gy))
                for t in ref[1:]:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM)
                    t /= world
                got = torch.autograd.grad(fused(x), [x] + list(fused.parameters()), gy)
                ok = ok and torch.equal(ref[0], got[0])  # grad_x stays local
                for a, b in zip(ref[1:], got[1:]):
                    # each rank's column sums are rounded to fp32 before the exchange: 1e-6 relative to the largest entry
                    ok = ok and (a - b).abs().max().item() <= 1e-6 * max(1.0, a.abs().max().item())
                    gathered = [torch.empty_like(b) for _ in range(world)]
                    dist.all_gather(gathered, b.contiguous())
                    ok = ok and all(torch.equal(gathered[0], t) for t in gathered)  # bitwise identical on every rank
            ok = ok and fused._grad_allreduce.status() == 0
        # inside a replayed CUDA graph
        torch.manual_seed(11)
        fused = QuantumConv1d(80, 384, 3, padding=1, n_qubits=4).to(dev).fuse_grad_allreduce()
        x = torch.full((2, 80, 256), float(rank + 1), device=dev) + torch.randn(2, 80, 256, device=dev)
        gy = torch.randn(2, 384, 256, device=dev)
        eager = torch.autograd.grad(fused(x), list(fused.parameters()), gy)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            captured = torch.autograd.grad(fused(x), list(fused.parameters()), gy)
        torch.cuda.synchronize()
        dist.barrier()
        for _ in range(3):
            graph.replay()
            torch.cuda.synchronize()
            ok = ok and all(torch.equal(a, b) for a, b in zip(eager, captured))
        ok = ok and fused._grad_allreduce.status() == 0
        # a rank that is LATE by more than a second (checkpointing, evaluation, a data-loader stall) must still get -- and give --
        # the true mean: the kernels wait like an NCCL collective (round 1 gave up after ~1 s and kept the local gradient)
        torch.manual_seed(11)
        plain = QuantumConv1d(80, 384, 3, padding=1, n_qubits=4).to(dev)
        ref = list(torch.autograd.grad(plain(x), list(plain.parameters()), gy))
        for t in ref:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t /= world
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 1:
            import time
            time.sleep(1.6)
        late = torch.autograd.grad(fused(x), list(fused.parameters()), gy)
        torch.cuda.synchronize()
        for a, b in zip(ref, late):
            ok = ok and (a - b).abs().max().item() <= 1e-6 * max(1.0, a.abs().max().item())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_fused_backward_allreduce_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_fused, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}


def _worker_stem(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import qasr_ijcnlp_b200 as qw
        from qasr_ijcnlp_b200 import stem_train_forward
        torch.manual_seed(5)  # same parameters on every rank
        c1 = qw.QuantumConv1d(80, 384, 3, padding=1, n_qubits=4).to(dev)
        c2 = qw.QuantumConv1d(384, 384, 3, stride=2, padding=1, n_qubits=4).to(dev)
        prm = list(c1.parameters()) + list(c2.parameters())
        x = torch.randn(2, 80, 640, device=dev, generator=torch.Generator(device=dev).manual_seed(10 + rank))  # a different shard per rank
        # reference: the helper without any collective, then NCCL mean
        ref = torch.autograd.grad(stem_train_forward(c1, c2, x, gelu=(True, False)).square().mean(), prm)
        ref = [g.clone() for g in ref]
        for g in ref:
            dist.all_reduce(g, op=dist.ReduceOp.AVG)
        # the layers' fused gradient all-reduce contexts handed to the chained backward by the helper
        c1.fuse_grad_allreduce()
        c2.fuse_grad_allreduce()
        ok = True
        for _ in range(3):  # several epochs of the exchange
            got = torch.autograd.grad(stem_train_forward(c1, c2, x, gelu=(True, False)).square().mean(), prm)
            for a, b in zip(got, ref):
                ok = ok and (a - b).abs().max().item() <= 1e-6 * max(1.0, b.abs().max().item())
        flat = torch.cat([g.reshape(-1) for g in got])
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        ok = ok and torch.equal(both[0], both[1])  # bitwise identical on the two ranks
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_stem_helper_with_fused_gradient_allreduce():
    """stem_train_forward on layers with fuse_grad_allreduce(): conv2's backward (no grad_x) through qw_conv1d_backward_dp, conv1's
    through qw_conv1d_backward_chained with the peer tables -- gradients equal the NCCL mean of the single-rank gradients."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_stem, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}
