"""Multi-process data-parallel logic on CPU (gloo, world_size 2): the gradient all-reduce that precedes each
optimizer step (b200gan/ddp.py) reproduces the single-process full-batch gradient for
networks without batch statistics, and all ranks end up with identical parameters after the step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model():
    torch.manual_seed(0)  # identical init on every rank (bench.py does the same)
    return torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.LeakyReLU(0.2), torch.nn.Linear(16, 1))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pytorch-gan_b200"))
    from b200gan import ddp
    net = _model()
    opt = torch.optim.Adam(net.parameters(), lr=2e-4, betas=(0.5, 0.999))
    g = torch.Generator().manual_seed(123)
    x_all, y_all = torch.randn(8, 12, generator=g), torch.randn(8, 1, generator=g)
    shard = slice(rank * 4, rank * 4 + 4)
    reducer = ddp.GradReducer(list(net.parameters()), world)
    loss = torch.nn.MSELoss()(net(x_all[shard]), y_all[shard])
    loss.backward()
    if rank == 0:
        reducer()        # all-reduce(sum) / world, right before the optimizer step (dcgan.py:169,183)
        grads = [p.grad.clone() for p in net.parameters()]
        opt.step()
    else:                # the form the training steps use (train._opt_step): reduce + step in one call
        from b200gan import train
        train._opt_step(opt, reducer)
        train._join(reducer)
        grads = [p.grad.clone() for p in net.parameters()]
    out[rank] = (grads, [p.detach().clone() for p in net.parameters()])
    dist.destroy_process_group()


def test_grad_reducer_world_size_2_matches_full_batch():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()  # no fork() of this multi-threaded process
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    net = _model()
    opt = torch.optim.Adam(net.parameters(), lr=2e-4, betas=(0.5, 0.999))
    g = torch.Generator().manual_seed(123)
    x_all, y_all = torch.randn(8, 12, generator=g), torch.randn(8, 1, generator=g)
    torch.nn.MSELoss()(net(x_all), y_all).backward()
    ref_grads = [p.grad.clone() for p in net.parameters()]
    opt.step()
    for r in range(world):
        grads, params = out[r]
        for a, b in zip(grads, ref_grads):     # mean of equal shards == full-batch mean
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
        for a, b in zip(params, net.parameters()):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-8)
    for a, b in zip(out[0][1], out[1][1]):     # replicas stay in lock-step
        assert torch.equal(a, b)
