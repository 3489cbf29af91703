"""GPU parity of the whole DCGAN path (BASELINE config 1) against the reference's golden vectors and
against the oracle restatement running stock torch on the same GPU."""
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


def _ours_from(g_ref, d_ref, img_size):
    from b200gan import zoo
    g = zoo.DCGANGenerator(img_size)
    d = zoo.DCGANDiscriminator(img_size)
    g.load_state_dict(g_ref.state_dict())
    d.load_state_dict(d_ref.state_dict())
    return g.cuda(), d.cuda()


def _no_dropout(m):
    for s in m.modules():
        if isinstance(s, torch.nn.Dropout2d):
            s.p = 0.0


def test_dcgan_against_reference_golden(golden_dir):
    """Inputs and expected outputs were produced by the UNMODIFIED reference (oracle/make_golden.py)."""
    fix = torch.load(os.path.join(golden_dir, "dcgan_32_b8.pt"), weights_only=False)
    g_cpu, d_cpu = ref_models.build_dcgan(fix["img_size"], seed=fix["seed"])
    g, d = _ours_from(g_cpu, d_cpu, fix["img_size"])
    _no_dropout(d)
    z, imgs = fix["z"].cuda(), fix["imgs"].cuda()
    gen = g(z)
    assert rel_err(gen, fix["gen"]) < TOL
    validity = d(gen)
    assert rel_err(validity, fix["validity"]) < TOL
    n = fix["batch"]
    bce = torch.nn.BCELoss()
    g_loss = bce(validity, torch.ones(n, 1, device="cuda"))
    assert abs(g_loss.item() - fix["g_loss"].item()) < TOL * abs(fix["g_loss"].item())
    g_loss.backward()
    for k, p in g.named_parameters():
        ref = fix["g_grads"][k]
        if ref["norm"] < 1e-7:  # conv bias in front of BatchNorm: analytically zero gradient
            continue
        assert abs(p.grad.double().norm().item() - ref["norm"]) < 2 * TOL * ref["norm"], k
        assert rel_err(p.grad.flatten()[:64], ref["head"]) < 5 * TOL, k
    for k, v in fix["bn_running"].items():
        assert rel_err(g.state_dict()[k].float(), v.float()) < TOL, k
    d.zero_grad()
    real_v = d(imgs)
    assert rel_err(real_v, fix["real_v"]) < TOL
    d_loss = (bce(real_v, torch.ones(n, 1, device="cuda")) + bce(d(gen.detach()), torch.zeros(n, 1, device="cuda"))) / 2
    assert abs(d_loss.item() - fix["d_loss"].item()) < TOL * abs(fix["d_loss"].item())
    d_loss.backward()
    for k, p in d.named_parameters():
        ref = fix["d_grads"][k]
        if ref["norm"] < 1e-7:
            continue
        assert abs(p.grad.double().norm().item() - ref["norm"]) < 2 * TOL * ref["norm"], k


@pytest.mark.parametrize("img_size,batch", [(64, 128), (32, 16)])
def test_dcgan_forward_backward_vs_stock_torch_on_gpu(img_size, batch):
    """BASELINE config 1 at full size: every output and parameter gradient of G and D."""
    g_ref, d_ref = ref_models.build_dcgan(img_size, seed=0)
    g, d = _ours_from(g_ref, d_ref, img_size)
    g_ref, d_ref = g_ref.cuda(), d_ref.cuda()
    for m in (d, d_ref):
        _no_dropout(m)
    z = ref_models.synthetic_z(batch, seed=1).cuda()
    imgs = ref_models.synthetic_images(batch, 1, img_size, img_size, seed=1).cuda()
    bce = torch.nn.BCELoss()
    ones = torch.ones(batch, 1, device="cuda")
    out = {}
    for tag, (gg, dd) in {"ref": (g_ref, d_ref), "ours": (g, d)}.items():
        gen = gg(z)
        v = dd(gen)
        loss = bce(v, ones) + bce(dd(imgs), ones * 0.9)
        loss.backward()
        out[tag] = (gen.detach(), v.detach(), loss.detach())
    assert out["ours"][0].shape == (batch, 1, img_size, img_size)
    assert rel_err(out["ours"][0], out["ref"][0]) < TOL
    assert rel_err(out["ours"][1], out["ref"][1]) < TOL
    assert rel_err(out["ours"][2], out["ref"][2]) < TOL
    for (k, po), (_, pr) in list(zip(g.named_parameters(), g_ref.named_parameters())) + \
            list(zip(d.named_parameters(), d_ref.named_parameters())):
        if pr.grad.double().norm().item() < 1e-7:
            continue
        assert rel_err(po.grad, pr.grad) < 3 * TOL, k
    for k, v in g_ref.state_dict().items():
        if "running" in k:
            assert rel_err(g.state_dict()[k], v) < TOL, k


def test_dcgan_training_steps_with_dropout_and_adam():
    """Three full steps (dcgan.py:146-183) with Dropout2d active: identical masks (same torch RNG calls),
    losses and post-Adam parameters.  Conv biases feeding a BatchNorm are excluded: their true gradient is
    zero and Adam amplifies rounding noise into +-lr steps."""
    from b200gan import train
    img_size, batch = 32, 32
    g_ref, d_ref = ref_models.build_dcgan(img_size, seed=0)
    g, d = _ours_from(g_ref, d_ref, img_size)
    g_ref, d_ref = g_ref.cuda(), d_ref.cuda()
    og_r, od_r = ref_models.make_adam(g_ref.parameters()), ref_models.make_adam(d_ref.parameters())
    og, od = ref_models.make_adam(g.parameters()), ref_models.make_adam(d.parameters())
    for step in range(3):
        z = ref_models.synthetic_z(batch, seed=10 + step).cuda()
        imgs = ref_models.synthetic_images(batch, 1, img_size, img_size, seed=10 + step).cuda()
        torch.manual_seed(100 + step)
        gl_r, dl_r, _ = ref_models.dcgan_step(g_ref, d_ref, og_r, od_r, imgs, z)
        torch.manual_seed(100 + step)
        gl, dl, _ = train.dcgan_step(g, d, og, od, imgs, z)
        assert abs(gl.item() - gl_r.item()) < 5 * TOL * abs(gl_r.item()), step
        assert abs(dl.item() - dl_r.item()) < 5 * TOL * abs(dl_r.item()), step
    skip = {"conv_blocks.2.bias", "conv_blocks.6.bias"}  # Conv -> BatchNorm directly (dcgan.py:55-56,59-60)
    for (k, po), (_, pr) in list(zip(g.named_parameters(), g_ref.named_parameters())) + \
            list(zip(d.named_parameters(), d_ref.named_parameters())):
        if k in skip:
            continue
        assert rel_err(po, pr) < 5 * TOL, k
