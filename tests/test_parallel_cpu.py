"""CPU tier: the data-parallel plumbing with gloo, world_size 2 (N>1 host logic without GPUs)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.GroupNorm(4, 8), nn.ELU(), nn.Conv2d(8, 1, 3, padding=1))


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from packnet_sfm_b200 import parallel
    torch.manual_seed(100 + rank)                      # deliberately different init per rank
    model = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.GroupNorm(4, 8), nn.ELU(), nn.Conv2d(8, 1, 3, padding=1))
    parallel.broadcast_parameters(model, src=0)
    bucket = parallel.FlatBucket(model.parameters())
    g = torch.Generator().manual_seed(7)
    x = torch.rand(4, 3, 8, 8, generator=g)
    sl = parallel.shard_batch(4, rank, world)
    bucket.zero_grad()
    model(x[sl]).pow(2).mean().backward()
    bucket.allreduce_mean()
    ret[rank] = (bucket.flat_grad.clone(), torch.cat([p.detach().reshape(-1) for p in model.parameters()]))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_equals_full_batch_gradient():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    g0, p0 = ret[0]
    g1, p1 = ret[1]
    assert torch.equal(p0, p1), "broadcast_parameters must make the replicas identical"
    assert torch.allclose(g0, g1, rtol=0, atol=0)
    # single-process gradient of the concatenated batch with rank 0's parameters
    torch.manual_seed(100)
    model = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.GroupNorm(4, 8), nn.ELU(), nn.Conv2d(8, 1, 3, padding=1))
    g = torch.Generator().manual_seed(7)
    x = torch.rand(4, 3, 8, 8, generator=g)
    # mean over the full batch == mean of the per-shard means (equal shard sizes)
    loss = 0.5 * (model(x[:2]).pow(2).mean() + model(x[2:]).pow(2).mean())
    loss.backward()
    full = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(g0, full, rtol=1e-5, atol=1e-7)


def test_flat_bucket_pack_gathers_fresh_gradients_into_views():
    from packnet_sfm_b200 import parallel
    m = _model()
    b = parallel.FlatBucket(m.parameters())
    for _ in range(2):
        b.zero_grad()
        assert all(p.grad is None for p in m.parameters())
        m(torch.rand(1, 3, 8, 8)).sum().backward()
        fresh = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone()
        b.pack()
        off = 0
        for p in m.parameters():
            assert p.grad.data_ptr() == b.flat_grad.data_ptr() + 4 * off
            off += p.numel()
        assert torch.equal(b.flat_grad, fresh)
    # a parameter that took no part in the backward packs as zeros
    b.zero_grad()
    b.pack()
    assert float(b.flat_grad.abs().sum()) == 0.0
    assert b.nbytes() == 4 * sum(p.numel() for p in m.parameters())
    assert b.allreduce_mean() is None      # no process group: a no-op


def _flat_adam_worker(rank, world, port, emu_so, ret):
    """optim.FlatAdam's bucketed all-reduce on gloo: stored weights (tiles written by the emulated pn_adam_step), two buckets
    launched from the 'backward' through grad_ready(), the tail (remaining stored weight + plain parameters) by
    allreduce_mean()."""
    import ctypes
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from packnet_sfm_b200 import _lib, _lib_conv, optim
    lib = ctypes.CDLL(emu_so)
    _lib._declare(lib)
    _lib_conv.declare(lib)
    _lib.lib = lambda: lib
    _lib.require_cuda = lambda *a: None
    _lib.current_stream = lambda: None
    torch.manual_seed(3)                                    # same parameters on both ranks
    ws = [torch.nn.Parameter(torch.rand(16, c, 3, 3) - 0.5) for c in (8, 70, 64, 5)]
    plain = [torch.nn.Parameter(torch.rand(n) - 0.5) for n in (16, 3, 2500)]
    opt = optim.FlatAdam(ws + plain, lr=1e-2, native=ws, buckets=[[ws[2]], [ws[1], ws[0]]])      # ws[3]: stored, in no bucket
    assert [r[2] for r in opt._bucket_ranges] == [1, 2] and opt._tail_start == opt._bucket_ranges[1][1]
    g = torch.Generator().manual_seed(50 + rank)            # different gradients per rank
    grads = [torch.rand(p.shape, generator=g) - 0.5 for p in ws + plain]
    opt.zero_grad()
    launched = []
    for i in (3, 2, 1, 0):                                  # "backward": the weight gradients arrive in reverse order
        ws[i].grad.copy_(grads[i])
        opt.grad_ready(ws[i]._pn_native)
        launched.append(list(opt._launched))
    for p, gr in zip(plain, grads[4:]):
        p.grad = gr.clone()
    opt.allreduce_mean()
    reduced = [p.grad.detach().clone() for p in ws + plain]
    opt.step()
    ret[rank] = (grads, reduced, launched, [p.detach().clone() for p in ws + plain])
    dist.destroy_process_group()


def test_flat_adam_buckets_allreduce_on_gloo(emu_lib):
    from conftest import EMU_SO
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_flat_adam_worker, args=(world, port, EMU_SO, ret), nprocs=world, join=True)
    g0, r0, l0, p0 = ret[0]
    g1, r1, l1, p1 = ret[1]
    # bucket 0 = {ws[2]} completes after the second gradient, bucket 1 = {ws[1], ws[0]} after the fourth
    assert l0 == [[False, False], [True, False], [True, False], [True, True]] == l1
    for a, b, ra, rb in zip(g0, g1, r0, r1):
        want = 0.5 * (a + b)
        assert torch.allclose(ra, want, rtol=0, atol=1e-7) and torch.equal(ra, rb)
    for a, b in zip(p0, p1):                                # identical replicas after the step
        assert torch.equal(a, b)
