"""GPU parity of the fused narrow conv chain (csrc/narrow_block.cu): runs of [Conv2d -> LeakyReLU -> Dropout2d] (+
BatchNorm2d) blocks (the DCGAN discriminator, dcgan.py:77-88) where the normalised tensors are never materialised, against
stock torch fp32 on the same GPU: outputs, the input gradient and every parameter gradient, running statistics.  All
arithmetic is fp32, so the bound is 1e-4 (re-association only)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _disc(ns, chans, k=3, p=0.25, last_bn=True):
    layers = []
    for i, (cin, cout) in enumerate(zip(chans[:-1], chans[1:])):
        layers += [ns.Conv2d(cin, cout, k, 2, 1), ns.LeakyReLU(0.2, inplace=True), ns.Dropout2d(p)]
        if i > 0 and (last_bn or i < len(chans) - 2):
            layers.append(ns.BatchNorm2d(cout, 0.8))
    return ns.Sequential(*layers)


def _pair(chans, **kw):
    from b200gan import zoo
    torch.manual_seed(5)
    ref = _disc(zoo.namespace(stock=True), chans, **kw).cuda().train()
    ours = _disc(zoo.namespace(), chans, **kw).cuda().train()
    with torch.no_grad():
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.normal_(1.0, 0.2)
                m.bias.normal_(0.0, 0.2)
    ours.load_state_dict(ref.state_dict())
    assert any(type(s).__name__ == "_ChainStep" for s in ours._plan())
    return ref, ours


def _check(ref, ours, x, seed=None, tol=1e-4, input_grad=True):
    xr, xo = x.clone().requires_grad_(input_grad), x.clone().requires_grad_(input_grad)
    if seed is not None:
        torch.manual_seed(seed)
    yr = ref(xr)
    if seed is not None:
        torch.manual_seed(seed)
    yo = ours(xo)
    assert yo.shape == yr.shape and yo.is_contiguous() == yr.is_contiguous()
    assert rel_err(yo, yr) < tol
    gy = torch.randn_like(yr)
    yr.backward(gy)
    yo.backward(gy)
    if input_grad:
        assert rel_err(xo.grad, xr.grad) < tol
    for (name, po), (_, pr) in zip(ours.named_parameters(), ref.named_parameters()):
        if pr.grad is None:
            assert po.grad is None, name
            continue
        assert rel_err(po.grad, pr.grad) < tol, name
    for (name, bo), (_, br) in zip(ours.named_buffers(), ref.named_buffers()):
        assert rel_err(bo.float(), br.float()) < tol, name


@pytest.mark.parametrize("chans,size,n", [((1, 16, 32, 64, 128), 32, 8), ((1, 16, 32, 64, 128), 64, 128),
                                          ((4, 8, 16), 20, 3), ((4, 16, 32, 64), 24, 5)])
def test_chain_without_dropout(chans, size, n):
    ref, ours = _pair(chans, p=0.0)
    _check(ref, ours, torch.randn(n, chans[0], size, size, device="cuda"))


def test_chain_with_dropout_masks_and_second_step():
    """Dropout2d active: the chain draws its masks with the same torch calls in the same order (seeded identically);
    two consecutive passes also check the running statistics and num_batches_tracked."""
    ref, ours = _pair((1, 16, 32, 64, 128))
    for step in range(2):
        ref.zero_grad(); ours.zero_grad()
        _check(ref, ours, torch.randn(16, 1, 64, 64, device="cuda"), seed=40 + step)


def test_chain_4x4_kernels_and_no_final_norm():
    ref, ours = _pair((4, 8, 16, 32), k=4, p=0.0, last_bn=False)
    _check(ref, ours, torch.randn(4, 4, 32, 32, device="cuda"))


def test_chain_with_frozen_weights_gives_the_input_gradient_only():
    """The generator step back-propagates through the discriminator with its parameters frozen (train.frozen)."""
    from b200gan import train
    ref, ours = _pair((1, 16, 32, 64, 128), p=0.0)
    x = torch.randn(8, 1, 32, 32, device="cuda")
    with train.frozen(ref), train.frozen(ours):
        _check(ref, ours, x)
    assert all(p.grad is None for p in ours.parameters())


def test_whole_discriminator_real_plus_fake_accumulates():
    """Two passes through the same weights in one backward (d_loss = (real + fake) / 2, dcgan.py:178-182)."""
    from b200gan import zoo
    torch.manual_seed(6)
    ref = zoo.DCGANDiscriminator(64, nn=zoo.namespace(stock=True)).cuda()
    ours = zoo.DCGANDiscriminator(64).cuda()
    ours.load_state_dict(ref.state_dict())
    for m in list(ref.modules()) + list(ours.modules()):
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0
    a, b = torch.randn(32, 1, 64, 64, device="cuda"), torch.randn(32, 1, 64, 64, device="cuda")
    bce = torch.nn.BCELoss()
    ones, zeros = torch.ones(32, 1, device="cuda"), torch.zeros(32, 1, device="cuda")
    for net in (ref, ours):
        ((bce(net(a), ones) + bce(net(b), zeros)) / 2).backward()
    for (name, po), (_, pr) in zip(ours.named_parameters(), ref.named_parameters()):
        assert rel_err(po.grad, pr.grad) < 1e-4, name


@pytest.mark.parametrize("chans,size,n,groups", [((1, 16, 32, 64, 128), 64, 128, 2), ((1, 16, 32, 64, 128), 32, 8, 2),
                                                 ((4, 16, 32), 16, 4, 4)])
def test_grouped_pass_equals_separate_passes_of_stock_torch(chans, size, n, groups):
    """ops.bn_groups(G): ONE pass of the fused chain over G concatenated batches against G separate forward passes of the
    stock modules (dcgan.py:178-179: discriminator(real_imgs), discriminator(gen_imgs.detach())) with Dropout2d active
    and the same seed: outputs, per-pass input gradients, parameter gradients summed over the passes, running
    statistics and num_batches_tracked after G sequential updates."""
    from b200gan import nn as bnn, ops
    ref, ours = _pair(chans)
    xs = [torch.randn(n, chans[0], size, size, device="cuda") for _ in range(groups)]
    assert bnn.groups_eligible(ours, (groups * n,) + tuple(xs[0].shape[1:]), groups)
    xr = [x.clone().requires_grad_(True) for x in xs]
    torch.manual_seed(77)
    yr = [ref(x) for x in xr]                       # pass g draws its masks after pass g-1 drew all of its own
    xo = torch.cat(xs).requires_grad_(True)
    torch.manual_seed(77)
    with ops.bn_groups(groups):
        yo = ours(xo)
    assert rel_err(yo, torch.cat(yr)) < 1e-4
    gy = [torch.randn_like(y) for y in yr]
    for y, g in zip(yr, gy):
        y.backward(g)
    yo.backward(torch.cat(gy))
    assert rel_err(xo.grad, torch.cat([x.grad for x in xr])) < 1e-4
    for (name, po), (_, pr) in zip(ours.named_parameters(), ref.named_parameters()):
        assert rel_err(po.grad, pr.grad) < 2e-4, name
    for (name, bo), (_, br) in zip(ours.named_buffers(), ref.named_buffers()):
        assert rel_err(bo.float(), br.float()) < 1e-4, name


def test_groups_refuse_modules_outside_a_fused_chain():
    from b200gan import nn as bnn, ops
    bn = bnn.BatchNorm2d(8).cuda().train()
    with ops.bn_groups(2), pytest.raises(RuntimeError):
        bn(torch.randn(4, 8, 8, 8, device="cuda"))
    seq = bnn.Sequential(bnn.Conv2d(256, 256, 3, 1, 1), bnn.BatchNorm2d(256)).cuda().train()   # too wide for the chain
    assert not bnn.groups_eligible(seq, (4, 256, 8, 8), 2)
