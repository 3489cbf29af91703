"""Kernel-level parity of the fused Discriminator chain (csrc/narrow_block.cu) against stock torch fp32 ops.

Reference: dcgan.py:77-88 -- Conv2d(in, out, 3, 2, 1) -> LeakyReLU(0.2) -> Dropout2d(0.25) [-> BatchNorm2d(out, 0.8)].  Every
kernel consumes the STORED output a of the previous block and applies that block's BatchNorm while loading, so the oracle
for each is the stock op on batch_norm(a):
    nb_fprop   y = dropout_scale * lrelu(conv2d(batch_norm(a), w, b)) and the batch sums of y
    nb_wgrad   conv2d_weight(batch_norm(a), dz)
    nb_dgrad   conv2d_input(dz, w) and the two sums BatchNorm's backward needs (sum G, sum G * ahat)
fp32 SIMT kernels: tolerance 1e-4 relative (summation order only)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(autouse=True)
def _fp32_reference():
    """The oracle ops must run in true fp32 (cuDNN would otherwise pick TF32 kernels)."""
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev

# (N, C, H, K, kernel, stride, pad): the four DCGAN blocks at the BASELINE size, odd sizes, 4x4 kernels, stride 1
GEOMS = [(128, 1, 64, 16, 3, 2, 1), (128, 16, 32, 32, 3, 2, 1), (128, 32, 16, 64, 3, 2, 1), (128, 64, 8, 128, 3, 2, 1),
         (3, 4, 20, 8, 3, 2, 1), (5, 16, 13, 32, 3, 2, 1), (6, 8, 24, 16, 4, 2, 1), (2, 32, 9, 64, 3, 1, 1),
         (7, 1, 31, 4, 3, 2, 1), (2, 64, 16, 128, 3, 2, 1)]


def _setup(geom, with_bn, seed=0):
    from b200gan import ops
    from b200gan.functional import PACK_SIMT_FPROP, PACK_SIMT_DGRAD
    n, c, h, k, ks, st, pad = geom
    gen = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(n, c, h, h, device="cuda", generator=gen).contiguous(memory_format=torch.channels_last)
    w = torch.randn(k, c, ks, ks, device="cuda", generator=gen) * 0.1
    b = torch.randn(k, device="cuda", generator=gen) * 0.1
    g, _ = ops.make_geom((n, c, h, h), (k, c, ks, ks), st, (pad,) * 4)
    edge, x = None, a
    if with_bn:
        gamma = torch.rand(c, device="cuda", generator=gen) + 0.5
        beta = torch.randn(c, device="cuda", generator=gen) * 0.1
        ad = a.double()
        stats = torch.cat([ad.sum((0, 2, 3)), (ad * ad).sum((0, 2, 3))]).contiguous()
        edge = ops.BnEdge(stats, gamma, beta, 0.8, n * h * h)
        x = F.batch_norm(a, None, None, gamma, beta, True, 0.0, 0.8)
    packs = (ops.pack_weights(g, w, PACK_SIMT_FPROP), ops.pack_weights(g, w, PACK_SIMT_DGRAD))
    return g, a, x, w, b, edge, packs


@pytest.mark.parametrize("with_bn", [False, True])
@pytest.mark.parametrize("geom", GEOMS)
def test_nb_fprop(geom, with_bn):
    from b200gan import ops
    from b200gan.functional import ACT_LRELU
    if geom[1] == 1 and with_bn:
        pytest.skip("the first block has no BatchNorm in front")
    g, a, x, w, b, edge, (pf, _) = _setup(geom, with_bn)
    n, c, h, k, ks, st, pad = geom
    cs = (torch.rand(n, k, device="cuda") > 0.25).float() / 0.75
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    nbt = torch.zeros((), device="cuda", dtype=torch.int64)
    y, stats = ops.nb_fprop(g, a, pf, b, ACT_LRELU, 0.2, cs, edge, rm if with_bn else None, rv if with_bn else None,
                            nbt if with_bn else None, 0.1, True)
    ref = F.leaky_relu(F.conv2d(x, w, b, st, pad), 0.2) * cs.view(n, k, 1, 1)
    assert rel_err(y, ref) < TOL
    rd = ref.double()
    assert rel_err(stats[:k], rd.sum((0, 2, 3))) < TOL
    assert rel_err(stats[k:], (rd * rd).sum((0, 2, 3))) < TOL
    if with_bn:
        assert rel_err(rm, 0.1 * a.mean((0, 2, 3))) < TOL
        assert rel_err(rv, 0.9 + 0.1 * a.var((0, 2, 3), unbiased=True)) < TOL
        assert int(nbt.item()) == 1


@pytest.mark.parametrize("with_bn", [False, True])
@pytest.mark.parametrize("geom", GEOMS)
def test_nb_wgrad_and_dgrad(geom, with_bn):
    from b200gan import ops
    if geom[1] == 1 and with_bn:
        pytest.skip("the first block has no BatchNorm in front")
    g, a, x, w, b, edge, (_, pd) = _setup(geom, with_bn, seed=1)
    n, c, h, k, ks, st, pad = geom
    dz = torch.randn(n, k, g.P, g.Q, device="cuda").contiguous(memory_format=torch.channels_last)
    dw = ops.nb_wgrad(g, a, dz, edge, tuple(w.shape))
    assert rel_err(dw, torch.nn.grad.conv2d_weight(x, tuple(w.shape), dz, st, pad)) < TOL
    gx, sums = ops.nb_dgrad(g, dz, pd, edge, a)
    ref = torch.nn.grad.conv2d_input((n, c, h, h), w, dz, st, pad)
    assert rel_err(gx, ref) < TOL
    if with_bn:
        mean = a.mean((0, 2, 3), keepdim=True)
        ahat = (a - mean) / torch.sqrt(a.var((0, 2, 3), unbiased=False, keepdim=True) + 0.8)
        assert rel_err(sums[:c], ref.double().sum((0, 2, 3))) < 10 * TOL
        assert rel_err(sums[c:], (ref.double() * ahat.double()).sum((0, 2, 3))) < 10 * TOL
    else:
        assert sums is None


def test_nb_wgrad_general_geometries_through_the_conv_entry():
    """b200gan_conv2d_wgrad routes narrow layers (few input or output channels) to the same kernel, including the 7x7
    reflection-padded stems of cyclegan/models.py:49-50,90 and pix2pix's 4x4 output layer (models.py:103)."""
    from b200gan import ops
    from b200gan.functional import ALGO_SIMT
    torch.manual_seed(2)
    for (n, c, h, k, ks, st, pad, reflect) in [(2, 3, 32, 64, 7, 1, 3, True), (2, 64, 32, 3, 7, 1, 3, True),
                                               (4, 128, 16, 3, 4, 1, 1, False), (8, 64, 64, 1, 3, 1, 1, False)]:
        x = torch.randn(n, c, h, h, device="cuda").contiguous(memory_format=torch.channels_last)
        g, _ = ops.make_geom((n, c, h, h), (k, c, ks, ks), st, (pad,) * 4, 1 if reflect else 0)
        dy = torch.randn(n, k, g.P, g.Q, device="cuda").contiguous(memory_format=torch.channels_last)
        dw, _ = ops.conv_wgrad(g, x, dy, (k, c, ks, ks), False, ALGO_SIMT)
        xp = F.pad(x, (pad,) * 4, mode="reflect") if reflect else x
        ref = torch.nn.grad.conv2d_weight(xp, (k, c, ks, ks), dy, st, 0 if reflect else pad)
        assert rel_err(dw, ref) < TOL, (n, c, h, k, ks)
