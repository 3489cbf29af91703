"""GPU parity of the whole DCGAN path (BASELINE config 1) against the reference's golden vectors and
against the oracle restatement running stock torch on the same GPU.

Tolerance policy (measured by tools/dcgan_diag.py, see profiles/r1_dcgan_precision.txt):
  * forward outputs (images, validity, losses): 1e-3 norm-relative vs fp32 -- the north-star bar;
  * gradients: back-propagation through BatchNorm (eps = 0.8, dcgan.py:56) subtracts large common
    components, which amplifies operand rounding.  The reference's OWN default GPU path (cuDNN with
    allow_tf32 = True) deviates from fp32 by 1-2e-2 on the Generator gradients.  So a gradient passes if
    it is within 2e-3 of fp32 OR within 1.5x the deviation of stock torch TF32 from fp32, measured in the
    same test on identical inputs ("parity with the reference's own PyTorch/cuDNN path");
  * with B200GAN_ALGO=simt (pure fp32 kernels) gradients must be within 5e-3 of fp32 outright
    (fp32 re-association noise through the same ill-conditioned chain reaches 1.6e-3 at batch 128).
"""
import copy
import os

import pytest
import torch

from conftest import rel_err
from oracle import ref_models

pytestmark = pytest.mark.gpu
TOL = 1e-3


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


def _set_tf32(on):
    torch.backends.cudnn.allow_tf32 = on
    torch.backends.cuda.matmul.allow_tf32 = on


def _fwd_bwd(g, d, z, imgs):
    _no_dropout(d)
    bce = torch.nn.BCELoss()
    ones = torch.ones(z.shape[0], 1, device=z.device)
    g.zero_grad()
    d.zero_grad()
    gen = g(z)
    v = d(gen)
    loss = bce(v, ones) + bce(d(imgs), ones * 0.9)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in list(g.named_parameters()) + list(d.named_parameters())}
    return dict(gen=gen.detach(), v=v.detach(), loss=loss.detach(), grads=grads)


def _check_grads(ours, fp32, tf32):
    import b200gan
    simt = b200gan.Config.algo == "simt"
    for k, ref in fp32["grads"].items():
        if ref.double().norm().item() < 1e-7:  # conv bias in front of BatchNorm: exactly-zero gradient
            continue
        e = rel_err(ours["grads"][k], ref)
        bound = 5e-3 if simt else max(2e-3, 1.5 * rel_err(tf32["grads"][k], ref))
        assert e < bound, f"{k}: rel err {e:.2e} exceeds {bound:.2e}"


@pytest.mark.parametrize("img_size,batch", [(64, 128), (32, 16)])
def test_dcgan_forward_backward_vs_stock_torch_on_gpu(img_size, batch):
    """BASELINE config 1 at full size: every output and parameter gradient of G and D."""
    g_cpu, d_cpu = ref_models.build_dcgan(img_size, seed=0)
    z = ref_models.synthetic_z(batch, seed=1).cuda()
    imgs = ref_models.synthetic_images(batch, 1, img_size, img_size, seed=1).cuda()
    _set_tf32(False)
    g_fp32 = copy.deepcopy(g_cpu).cuda()
    fp32 = _fwd_bwd(g_fp32, copy.deepcopy(d_cpu).cuda(), z, imgs)
    _set_tf32(True)   # the reference's default GPU path (torch.backends.cudnn.allow_tf32 defaults to True)
    tf32 = _fwd_bwd(copy.deepcopy(g_cpu).cuda(), copy.deepcopy(d_cpu).cuda(), z, imgs)
    _set_tf32(False)
    g, d = _ours_from(g_cpu, d_cpu, img_size)
    ours = _fwd_bwd(g, d, z, imgs)
    assert ours["gen"].shape == (batch, 1, img_size, img_size)
    assert rel_err(ours["gen"], fp32["gen"]) < TOL
    assert rel_err(ours["v"], fp32["v"]) < TOL
    assert rel_err(ours["loss"], fp32["loss"]) < TOL
    _check_grads(ours, fp32, tf32)
    for k, v in g_fp32.state_dict().items():
        if "running" in k:  # running_var uses the unbiased variance
            assert rel_err(g.state_dict()[k].float(), v.float()) < TOL, k


def test_dcgan_against_reference_golden(golden_dir):
    """Inputs and expected outputs were produced by the UNMODIFIED reference on CPU (oracle/make_golden.py)."""
    import b200gan
    fix = torch.load(os.path.join(golden_dir, "dcgan_32_b8.pt"), weights_only=False)
    g_cpu, d_cpu = ref_models.build_dcgan(fix["img_size"], seed=fix["seed"])
    g, d = _ours_from(g_cpu, d_cpu, fix["img_size"])
    _no_dropout(d)
    z, imgs = fix["z"].cuda(), fix["imgs"].cuda()
    n = fix["batch"]
    bce = torch.nn.BCELoss()
    _set_tf32(False)
    gen = g(z)
    assert rel_err(gen, fix["gen"]) < TOL
    validity = d(gen)
    assert rel_err(validity, fix["validity"]) < TOL
    g_loss = bce(validity, torch.ones(n, 1, device="cuda"))
    assert abs(g_loss.item() - fix["g_loss"].item()) < TOL * abs(fix["g_loss"].item())
    g_loss.backward()
    # yardstick for the gradients: stock torch TF32 on this GPU vs the golden (reference, CPU fp32) values
    _set_tf32(True)
    g_t, d_t = copy.deepcopy(g_cpu).cuda(), copy.deepcopy(d_cpu).cuda()
    _no_dropout(d_t)
    bce(d_t(g_t(z)), torch.ones(n, 1, device="cuda")).backward()
    _set_tf32(False)
    for (k, p), (_, pt) in zip(g.named_parameters(), g_t.named_parameters()):
        ref = fix["g_grads"][k]
        if ref["norm"] < 1e-7:
            continue
        head = ref["head"].cuda()
        e = rel_err(p.grad.flatten()[:64], head)
        bound = 5e-3 if b200gan.Config.algo == "simt" else max(5e-3, 1.5 * rel_err(pt.grad.flatten()[:64], head))
        assert e < bound, f"{k}: {e:.2e} vs {bound:.2e}"
    for k, v in fix["bn_running"].items():
        assert rel_err(g.state_dict()[k].float(), v.float()) < TOL, k
    d.zero_grad()
    real_v = d(imgs)
    assert rel_err(real_v, fix["real_v"]) < TOL
    d_loss = (bce(real_v, torch.ones(n, 1, device="cuda")) + bce(d(gen.detach()), torch.zeros(n, 1, device="cuda"))) / 2
    assert abs(d_loss.item() - fix["d_loss"].item()) < TOL * abs(fix["d_loss"].item())
    d_loss.backward()
    # yardstick for the D gradients (conv3/conv4 of D run in TF32 on the tensor cores as well)
    _set_tf32(True)
    d_t.zero_grad()
    with torch.no_grad():
        gen_t = g_t(z)
    ((bce(d_t(imgs), torch.ones(n, 1, device="cuda")) + bce(d_t(gen_t), torch.zeros(n, 1, device="cuda"))) / 2).backward()
    _set_tf32(False)
    for (k, p), (_, pt) in zip(d.named_parameters(), d_t.named_parameters()):
        ref = fix["d_grads"][k]
        if ref["norm"] < 1e-7:
            continue
        e = abs(p.grad.double().norm().item() - ref["norm"]) / ref["norm"]
        e_t = abs(pt.grad.double().norm().item() - ref["norm"]) / ref["norm"]
        bound = 2 * TOL if b200gan.Config.algo == "simt" else max(5 * TOL, 1.5 * e_t)
        assert e < bound, f"{k}: {e:.2e} vs {bound:.2e}"


def test_dcgan_training_steps_with_dropout_and_adam():
    """Three full steps (dcgan.py:146-183) with Dropout2d active: identical masks (same torch RNG calls) and
    losses within 1e-3 of stock torch fp32 at every step (losses are forward quantities of the updated nets).
    Post-Adam parameters: Adam's m/sqrt(v) normalisation turns a gradient deviation of 1e-2 into a 1e-2
    deviation of a 2e-4 step, so parameters with a non-zero initial value stay within 1e-3 of the fp32 run;
    zero-initialised ones are compared through their accumulated update."""
    from b200gan import train
    _set_tf32(False)
    img_size, batch = 32, 32
    g_ref, d_ref = ref_models.build_dcgan(img_size, seed=0)
    g, d = _ours_from(g_ref, d_ref, img_size)
    g_ref, d_ref = g_ref.cuda(), d_ref.cuda()
    og_r, od_r = ref_models.make_adam(g_ref.parameters()), ref_models.make_adam(d_ref.parameters())
    og, od = ref_models.make_adam(g.parameters()), ref_models.make_adam(d.parameters())
    init_params = [p.detach().clone() for p in list(g_ref.parameters()) + list(d_ref.parameters())]
    # yardstick: the same three steps on stock torch with TF32 convolutions (the reference's default GPU path)
    g_t, d_t = copy.deepcopy(g_ref), copy.deepcopy(d_ref)
    og_t, od_t = ref_models.make_adam(g_t.parameters()), ref_models.make_adam(d_t.parameters())
    for step in range(3):
        z = ref_models.synthetic_z(batch, seed=10 + step).cuda()
        imgs = ref_models.synthetic_images(batch, 1, img_size, img_size, seed=10 + step).cuda()
        torch.manual_seed(100 + step)
        gl_r, dl_r, gen_r = ref_models.dcgan_step(g_ref, d_ref, og_r, od_r, imgs, z)
        _set_tf32(True)
        torch.manual_seed(100 + step)
        ref_models.dcgan_step(g_t, d_t, og_t, od_t, imgs, z)
        _set_tf32(False)
        torch.manual_seed(100 + step)
        gl, dl, gen = train.dcgan_step(g, d, og, od, imgs, z)
        assert abs(gl.item() - gl_r.item()) < TOL * abs(gl_r.item()), step
        assert abs(dl.item() - dl_r.item()) < TOL * abs(dl_r.item()), step
        assert rel_err(gen, gen_r) < 5 * TOL, step   # G parameters already carry 1-2 steps of TF32-level gradient noise
    import b200gan
    skip = {"conv_blocks.2.bias", "conv_blocks.6.bias"}  # Conv -> BatchNorm directly (dcgan.py:55-56,59-60)
    for (k, po), (_, pr), (_, pt), p0 in zip(list(g.named_parameters()) + list(d.named_parameters()),
                                             list(g_ref.named_parameters()) + list(d_ref.named_parameters()),
                                             list(g_t.named_parameters()) + list(d_t.named_parameters()),
                                             init_params):
        if k in skip:
            continue
        if p0.abs().max().item() == 0.0:
            # zero-initialised BatchNorm biases: the value after 3 steps IS the sum of the Adam steps, so it carries
            # the gradient-level deviation (TF32 through BatchNorm backward).  Bound: 1.5x what stock TF32 shows here.
            step_tol = 2e-2 if b200gan.Config.algo == "simt" else max(2e-2, 1.5 * rel_err(pt - p0, pr - p0))
            assert rel_err(po - p0, pr - p0) < step_tol, (k, step_tol)
        else:
            # three Adam steps on TF32-level gradient noise: 2e-3, or 1.5x what stock TF32 shows on this parameter
            assert rel_err(po, pr) < max(2 * TOL, 1.5 * rel_err(pt, pr)), k


def test_a_script_in_the_reference_idiom_runs_under_the_launcher_on_the_gpu():
    """launch.run() of a stand-alone script written in the reference's API idiom (tests/scripts/mini_convgan): torch.nn
    looked up by attribute, Sequential(*layers), .apply init by class name, torch.cuda.FloatTensor(numpy), Variable,
    torch.optim.Adam -- stock torch vs the b200gan drop-ins (+ the one-launch Adam), same seeds: printed losses agree."""
    from b200gan import launch
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts", "mini_convgan", "mini_convgan.py")
    args = ["--epochs", "1", "--batch_size", "16", "--side", "32"]
    _set_tf32(False)
    ref = launch.run(script, args, iters=3, seed=3, stock=True, quiet=True)
    ours = launch.run(script, args, iters=3, seed=3, stock=False, quiet=True)
    assert type(ours["opt_g"]).__module__.endswith("b200gan.optim")
    assert any(type(s).__name__ == "_ChainStep" for s in ours["D"].body._plan())
    assert len(ref["history"]) == 3 and len(ours["history"]) == 3
    for (dr, gr), (do, go) in zip(ref["history"], ours["history"]):
        assert abs(do - dr) < 2e-3 * abs(dr) and abs(go - gr) < 2e-3 * abs(gr), (ref["history"], ours["history"])
