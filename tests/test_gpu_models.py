"""GPU parity for BASELINE configs 3 and 4: the pix2pix U-Net / PatchGAN and the CycleGAN ResNet generator /
discriminator built from the drop-in modules, against (a) golden vectors produced by the reference's own
models.py on CPU (oracle/make_golden.py) and (b) the oracle restatement on stock torch fp32 on this GPU."""
import os

import pytest
import torch

from conftest import rel_err
from oracle import ref_models

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(autouse=True)
def _fp32_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _grad_norm_check(model, ref_norms, tol, what, yard=None):
    """Parameter gradient norms against the golden fp32 norms.  yard: the same model run by stock torch with TF32
    convolutions (the reference's default GPU arithmetic) -- where given, a parameter may deviate 1.5x as far as
    stock TF32 does on it (bias gradients are sums with heavy cancellation: theirs moves by 1e-2 too)."""
    top = max(ref_norms.values())
    yard_grads = dict((k, p.grad) for k, p in yard.named_parameters()) if yard is not None else {}
    for k, p in model.named_parameters():
        rn = ref_norms[k]
        if rn < 1e-4 * top:   # conv bias in front of InstanceNorm: exactly-zero gradient, only fp noise
            continue
        bound = tol
        if k in yard_grads and yard_grads[k] is not None:
            bound = max(tol, 2.0 * abs(yard_grads[k].double().norm().item() - rn) / rn)
        assert abs(p.grad.double().norm().item() - rn) < bound * rn, f"{what} {k} (bound {bound:.2e})"


def _set_tf32(on):
    torch.backends.cudnn.allow_tf32 = on
    torch.backends.cuda.matmul.allow_tf32 = on


def _bound(ref_fp32, ref_tf32):
    """1e-3, or 1.5x the deviation of the reference's own default GPU path (stock torch, TF32 convolutions) from
    fp32 on the same inputs -- a dozen TF32 convolutions in sequence exceed 1e-3 for stock torch as well."""
    return max(TOL, 1.5 * rel_err(ref_tf32, ref_fp32))


def test_pix2pix_against_reference_golden(golden_dir):
    from b200gan import zoo
    fix = torch.load(os.path.join(golden_dir, "pix2pix_256_b1.pt"), weights_only=False)
    g_cpu, d_cpu = ref_models.build_pix2pix(fix["seed"])
    g, d = zoo.GeneratorUNet(), zoo.Pix2PixDiscriminator()
    g.load_state_dict(g_cpu.state_dict())
    d.load_state_dict(d_cpu.state_dict())
    g, d = g.cuda().eval(), d.cuda().train()
    size, n = fix["size"], fix["batch"]
    real_a = ref_models.synthetic_images(n, 3, size, size, seed=fix["seed"] + 1).cuda()
    real_b = ref_models.synthetic_images(n, 3, size, size, seed=fix["seed"] + 2).cuda()
    fake_b = g(real_a)
    assert fake_b.shape == (n, 3, size, size)
    assert rel_err(fake_b[..., ::8, ::8], fix["fake_b"]) < TOL
    pred = d(fake_b, real_a)
    assert pred.shape == fix["pred"].shape
    assert rel_err(pred, fix["pred"]) < 2 * TOL
    loss = torch.nn.MSELoss()(pred, torch.ones_like(pred)) + 100 * torch.nn.L1Loss()(fake_b, real_b)
    assert abs(loss.item() - fix["loss"].item()) < TOL * abs(fix["loss"].item())
    loss.backward()
    # yardstick: the reference modules themselves on this GPU with TF32 convolutions
    g_t, d_t = g_cpu.cuda().eval(), d_cpu.cuda().train()
    _set_tf32(True)
    fake_t = g_t(real_a)
    pred_t = d_t(fake_t, real_a)
    (torch.nn.MSELoss()(pred_t, torch.ones_like(pred_t)) + 100 * torch.nn.L1Loss()(fake_t, real_b)).backward()
    _set_tf32(False)
    # batch 1: the U-Net bottleneck is 512 channels at 1x1 .. 4x4 pixels, where one TF32-flipped ReLU mask moves a whole
    # filter's gradient (stock TF32 itself is 0.7 % off on down8): 1.5e-2, or twice the stock TF32 deviation
    _grad_norm_check(g, fix["g_grad_norms"], 1.5e-2, "pix2pix G", g_t)
    _grad_norm_check(d, fix["d_grad_norms"], 1.5e-2, "pix2pix D", d_t)


def test_pix2pix_channels_last_chain_and_dropout_training_mode():
    """Training-mode forward (Dropout active) keeps every block channels_last (no layout round trips) and
    produces the statistics of dropout: finite outputs, right shapes, backward runs."""
    from b200gan import zoo
    torch.manual_seed(0)
    g = zoo.GeneratorUNet().cuda().train()
    x = torch.randn(2, 3, 256, 256, device="cuda")
    d1 = g.down1(x)
    assert d1.is_contiguous(memory_format=torch.channels_last) and not d1.is_contiguous()
    y = g(x)
    assert y.shape == (2, 3, 256, 256) and bool(torch.isfinite(y).all())
    y.mean().backward()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in g.parameters())


def test_cyclegan_against_reference_golden(golden_dir):
    from b200gan import zoo
    fix = torch.load(os.path.join(golden_dir, "cyclegan_64_b2.pt"), weights_only=False)
    size, n, blocks = fix["size"], fix["batch"], fix["blocks"]
    shape = (3, size, size)
    nets_cpu = ref_models.build_cyclegan(shape, blocks, fix["seed"])
    nets = [zoo.GeneratorResNet(shape, blocks), zoo.GeneratorResNet(shape, blocks), zoo.CycleGANDiscriminator(shape),
            zoo.CycleGANDiscriminator(shape)]
    for m, r in zip(nets, nets_cpu):
        m.load_state_dict(r.state_dict())
        m.cuda().train()
    g_ab, g_ba, d_a, d_b = nets
    real_a = ref_models.synthetic_images(n, 3, size, size, seed=fix["seed"] + 1).cuda()
    real_b = ref_models.synthetic_images(n, 3, size, size, seed=fix["seed"] + 2).cuda()
    mse, l1 = torch.nn.MSELoss(), torch.nn.L1Loss()
    valid = torch.ones(n, *d_a.output_shape, device="cuda")
    loss_id = (l1(g_ba(real_a), real_a) + l1(g_ab(real_b), real_b)) / 2
    fake_b, fake_a = g_ab(real_a), g_ba(real_b)
    _set_tf32(True)
    with torch.no_grad():
        tf32_b = nets_cpu[0].cuda()(real_a)[..., ::4, ::4]
    _set_tf32(False)
    bound = _bound(fix["fake_b"], tf32_b)
    assert rel_err(fake_b[..., ::4, ::4], fix["fake_b"]) < bound
    assert rel_err(fake_a[..., ::4, ::4], fix["fake_a"]) < bound
    loss_gan = (mse(d_b(fake_b), valid) + mse(d_a(fake_a), valid)) / 2
    loss_cyc = (l1(g_ba(fake_b), real_a) + l1(g_ab(fake_a), real_b)) / 2
    for ours, ref in zip((loss_id, loss_gan, loss_cyc), fix["parts"]):
        assert abs(ours.item() - ref) < 2 * bound * abs(ref)
    loss_g = loss_gan + 10.0 * loss_cyc + 5.0 * loss_id
    assert abs(loss_g.item() - fix["loss_g"].item()) < 2 * bound * abs(fix["loss_g"].item())
    loss_g.backward()
    _grad_norm_check(g_ab, fix["g_ab_grad_norms"], 2e-2, "cyclegan G_AB")
    _grad_norm_check(g_ba, fix["g_ba_grad_norms"], 2e-2, "cyclegan G_BA")


@pytest.mark.parametrize("which", ["pix2pix_d", "cyclegan_g"])
def test_models_vs_stock_torch_on_gpu(which):
    """Same weights, same inputs, stock torch fp32 on this GPU: outputs and all parameter gradients."""
    from b200gan import zoo
    if which == "pix2pix_d":
        _, ref = ref_models.build_pix2pix(1)
        ours = zoo.Pix2PixDiscriminator()
        args = [torch.randn(4, 3, 128, 128, device="cuda"), torch.randn(4, 3, 128, 128, device="cuda")]
    else:
        ref = ref_models.build_cyclegan((3, 64, 64), 3, 1)[0]
        ours = zoo.GeneratorResNet((3, 64, 64), 3)
        args = [torch.randn(2, 3, 64, 64, device="cuda")]
    ours.load_state_dict(ref.state_dict())
    ref, ours = ref.cuda().train(), ours.cuda().train()
    yr = ref(*args)
    yo = ours(*args)
    _set_tf32(True)
    with torch.no_grad():
        y_tf32 = ref(*args)
    _set_tf32(False)
    assert rel_err(yo, yr) < _bound(yr, y_tf32)
    gy = torch.randn_like(yr)
    yr.backward(gy)
    yo.backward(gy)
    # yardstick: the same backward on stock torch with TF32 convolutions (the reference's default GPU path)
    import copy
    ref_t = copy.deepcopy(ref)
    ref_t.zero_grad()
    _set_tf32(True)
    ref_t(*args).backward(gy)
    _set_tf32(False)
    for (k, po), (_, pr), (_, pt) in zip(ours.named_parameters(), ref.named_parameters(), ref_t.named_parameters()):
        if pr.grad.double().norm().item() < 1e-6 or k.endswith("bias"):
            continue
        bound = max(1e-2, 1.5 * rel_err(pt.grad, pr.grad))
        assert rel_err(po.grad, pr.grad) < bound, f"{k}: bound {bound:.2e}"


def test_pix2pix_and_cyclegan_steps_match_stock_torch():
    """One full training step of each script body (pix2pix.py:131-172, cyclegan.py:163-241) with the drop-in modules
    vs the same step on the oracle models (stock torch fp32): losses after the step's forward passes."""
    import itertools
    from b200gan import train, zoo
    adam = lambda ps: torch.optim.Adam(ps, lr=2e-4, betas=(0.5, 0.999))  # noqa: E731
    # pix2pix, batch 1 (Dropout active in the reference too -> compare in eval mode for G)
    g_ref, d_ref = ref_models.build_pix2pix(2)
    g, d = zoo.GeneratorUNet(), zoo.Pix2PixDiscriminator()
    g.load_state_dict(g_ref.state_dict()); d.load_state_dict(d_ref.state_dict())
    g_ref, d_ref, g, d = g_ref.cuda().eval(), d_ref.cuda(), g.cuda().eval(), d.cuda()
    a = ref_models.synthetic_images(1, 3, 256, 256, seed=5).cuda()
    b = ref_models.synthetic_images(1, 3, 256, 256, seed=6).cuda()
    lg_r, ld_r = train.pix2pix_step(g_ref, d_ref, adam(g_ref.parameters()), adam(d_ref.parameters()), a, b)
    lg, ld = train.pix2pix_step(g, d, adam(g.parameters()), adam(d.parameters()), a, b)
    assert abs(lg.item() - lg_r.item()) < 2 * TOL * abs(lg_r.item())
    assert abs(ld.item() - ld_r.item()) < 2 * TOL * abs(ld_r.item())
    # cyclegan, 64x64, 3 residual blocks, replay buffers seeded identically
    import random
    shape = (3, 64, 64)
    refs = [m.cuda() for m in ref_models.build_cyclegan(shape, 3, 4)]
    ours = [zoo.GeneratorResNet(shape, 3), zoo.GeneratorResNet(shape, 3), zoo.CycleGANDiscriminator(shape),
            zoo.CycleGANDiscriminator(shape)]
    for m, r in zip(ours, refs):
        m.load_state_dict(r.state_dict())
        m.cuda()
    a = ref_models.synthetic_images(2, 3, 64, 64, seed=7).cuda()
    b = ref_models.synthetic_images(2, 3, 64, 64, seed=8).cuda()
    res = []
    for nets in (refs, ours):
        random.seed(0)
        og = adam(itertools.chain(nets[0].parameters(), nets[1].parameters()))
        oa, ob = adam(nets[2].parameters()), adam(nets[3].parameters())
        ba, bb = train.ReplayBuffer(), train.ReplayBuffer()
        for _ in range(2):
            out = train.cyclegan_step(*nets, og, oa, ob, a, b, ba, bb)
        res.append(out)
    assert abs(res[1][0].item() - res[0][0].item()) < 5 * TOL * abs(res[0][0].item())
    assert abs(res[1][1].item() - res[0][1].item()) < 5 * TOL * abs(res[0][1].item())
