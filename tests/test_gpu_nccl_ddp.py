"""Data-parallel correctness on real GPUs (parity gap 4): the DCGAN step on 2 ranks over NCCL (b200gan/ddp.py:
flat bucket, all-reduce on a side stream overlapping the discriminator phase, 1/world folded into the Adam kernel)
equals the single-process emulation of the same thing -- two replicas with local BatchNorm statistics, gradients
averaged by hand, one optimizer step on the average.  Needs 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IMG, PER_RANK, STEPS = 32, 16, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(dev):
    from b200gan import zoo
    torch.manual_seed(0)
    g, d = zoo.DCGANGenerator(IMG).to(dev), zoo.DCGANDiscriminator(IMG).to(dev)
    g.apply(zoo.weights_init_normal)
    d.apply(zoo.weights_init_normal)
    for m in d.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0  # per-rank dropout streams are not what this test is about
    return g, d


def _data(step, rank):
    gen = torch.Generator().manual_seed(1000 + 10 * step + rank)
    return torch.rand(PER_RANK, 1, IMG, IMG, generator=gen) * 2 - 1, torch.randn(PER_RANK, 100, generator=gen)


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "pytorch-gan_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from b200gan import ddp, optim, train
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    g, d = _build(dev)
    og, od = optim.Adam(g.parameters(), lr=2e-4, betas=(0.5, 0.999)), optim.Adam(d.parameters(), lr=2e-4, betas=(0.5, 0.999))
    rg, rd = ddp.GradReducer(list(g.parameters()), world, og), ddp.GradReducer(list(d.parameters()), world, od)
    losses, local_grads = [], None
    for step in range(STEPS):
        imgs, z = _data(step, rank)
        gl, dl, _ = train.dcgan_step(g, d, og, od, imgs.to(dev), z.to(dev), reduce_g=rg, reduce_d=rd)
        losses.append((gl.item(), dl.item()))
        if step == 0:   # this rank's own (un-reduced) gradients of the first step: p.grad is left untouched by the reducer
            torch.cuda.synchronize()
            local_grads = {"g": {k: p.grad.detach().cpu() for k, p in g.named_parameters() if p.grad is not None},
                           "d": {k: p.grad.detach().cpu() for k, p in d.named_parameters() if p.grad is not None}}
    torch.cuda.synchronize()
    out[rank] = dict(losses=losses, local_grads=local_grads, g={k: v.cpu() for k, v in g.state_dict().items()},
                     d={k: v.cpu() for k, v in d.state_dict().items()})
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_dcgan_step_on_two_ranks_equals_single_process_emulation():
    import torch.multiprocessing as mp
    from b200gan import train
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)

    dev = torch.device("cuda", 0)
    reps = [_build(dev) for _ in range(world)]
    adam = lambda ps: torch.optim.Adam(ps, lr=2e-4, betas=(0.5, 0.999))  # noqa: E731
    og, od = adam(reps[0][0].parameters()), adam(reps[0][1].parameters())
    bce = torch.nn.BCELoss()
    ref_losses = [[] for _ in range(world)]
    ref_local = [{"g": {}, "d": {}} for _ in range(world)]
    zero_init = {key: {k for k, v in net.state_dict().items() if v.dtype.is_floating_point and float(v.abs().max()) == 0.0}
                 for key, net in (("g", reps[0][0]), ("d", reps[0][1]))}

    def average_into_first(nets):
        for ps in zip(*[list(n.parameters()) for n in nets]):
            ps[0].grad = sum(p.grad for p in ps) / world

    def broadcast_from_first(nets):
        with torch.no_grad():
            for ps in zip(*[list(n.parameters()) for n in nets]):
                for p in ps[1:]:
                    p.copy_(ps[0])

    for step in range(STEPS):
        batches = [tuple(t.to(dev) for t in _data(step, r)) for r in range(world)]
        ones, zeros = torch.ones(PER_RANK, 1, device=dev), torch.zeros(PER_RANK, 1, device=dev)
        gens = []
        for (g, d), (imgs, z) in zip(reps, batches):          # dcgan.py:157-168 on every replica
            g.zero_grad(set_to_none=True)
            gen = g(z)
            with train.frozen(d):
                gl = bce(d(gen), ones)
                gl.backward()
            gens.append((gen.detach(), gl.item()))
        if step == 0:
            for r, (g, _) in enumerate(reps):
                ref_local[r]["g"] = {k: p.grad.detach().clone() for k, p in g.named_parameters() if p.grad is not None}
        average_into_first([g for g, _ in reps])
        og.step()
        broadcast_from_first([g for g, _ in reps])
        for r, ((g, d), (imgs, z)) in enumerate(zip(reps, batches)):   # dcgan.py:175-182
            d.zero_grad(set_to_none=True)
            dl = (bce(d(imgs), ones) + bce(d(gens[r][0]), zeros)) / 2
            dl.backward()
            ref_losses[r].append((gens[r][1], dl.item()))
        if step == 0:
            for r, (_, d) in enumerate(reps):
                ref_local[r]["d"] = {k: p.grad.detach().clone() for k, p in d.named_parameters() if p.grad is not None}
        average_into_first([d for _, d in reps])
        od.step()
        broadcast_from_first([d for _, d in reps])

    worst_grad, worst_param, worst_zero = (0.0, None), (0.0, None), (0.0, None)
    for r in range(world):
        res = out[r]
        for (gl, dl), (gl_r, dl_r) in zip(res["losses"], ref_losses[r]):
            assert abs(gl - gl_r) < 1e-4 * abs(gl_r) and abs(dl - dl_r) < 1e-4 * abs(dl_r), (r, gl, gl_r, dl, dl_r)
        # (1) the rank's own first-step gradients: same kernels on the same data in both runs -- only the order of
        # floating-point atomics differs
        for key in ("g", "d"):
            top = max(v.double().norm().item() for v in ref_local[r][key].values())
            for k, v in ref_local[r][key].items():
                if v.double().norm().item() < 1e-4 * top:
                    continue   # conv bias in front of a BatchNorm: mathematically zero gradient, only rounding noise
                got = res["local_grads"][key][k].to(dev).double()
                e = (got - v.double()).norm().item() / (v.double().norm().item() or 1.0)
                if e > worst_grad[0]:
                    worst_grad = (e, (r, key, k))
        # (2) parameters after STEPS optimizer steps.  Zero-initialised ones (BatchNorm biases) ARE the sum of the Adam
        # updates lr * m / sqrt(v), which turn a relative gradient difference eps into an O(eps) difference of the whole
        # value; the others start at O(2e-2) and move by 2e-4 per step, which dilutes the same difference 100x.
        for net, key in ((reps[r][0], "g"), (reps[r][1], "d")):
            for k, v in net.state_dict().items():
                got = res[key][k].to(dev)
                if v.dtype.is_floating_point:
                    den = v.double().norm().item() or 1.0
                    e = (got.double() - v.double()).norm().item() / den
                    if k in zero_init[key]:
                        if e > worst_zero[0]:
                            worst_zero = (e, (r, key, k))
                    elif e > worst_param[0]:
                        worst_param = (e, (r, key, k))
                else:
                    assert torch.equal(got, v), (r, key, k)
    print(f"2-rank NCCL vs emulation: worst local gradient {worst_grad}, worst parameter {worst_param}, "
          f"worst zero-initialised parameter {worst_zero}")
    assert worst_grad[0] < 1e-4, worst_grad
    assert worst_param[0] < 2e-4, worst_param
    assert worst_zero[0] < 5e-3, worst_zero
    # replicas hold identical parameters (BatchNorm running statistics are per replica by design)
    for k, v in out[0]["g"].items():
        if "running" not in k and "num_batches" not in k:
            assert torch.equal(v, out[1]["g"][k]), k
