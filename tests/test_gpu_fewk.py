"""Stride-1 convolutions with K <= 4 output channels (csrc/fewk.cu) against stock torch fp32, through the C-ABI conv entry
points with the SIMT algorithm.

Reference layers:
    pix2pix/models.py:97-102    Upsample(x2) -> ZeroPad2d((1, 0, 1, 0)) -> Conv2d(128, 3, 4, padding=1) -> Tanh
    pix2pix/models.py:131-132   ZeroPad2d((1, 0, 1, 0)) -> Conv2d(512, 1, 4, padding=1, bias=False)
    cyclegan/models.py:88-90    ReflectionPad2d(3) -> Conv2d(64, 3, 7) -> Tanh
fp32 FMA kernels: 1e-4 relative (summation order only)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(autouse=True)
def _fp32_reference():
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


# (N, C, H, W, K, kernel, pads (top, left, bottom, right), up, reflect)
CASES = [(2, 128, 16, 16, 3, 4, (2, 2, 1, 1), 2, False),     # pix2pix generator output layer
         (3, 128, 9, 13, 3, 4, (2, 2, 1, 1), 2, False),      # ragged: tiles hang over both edges
         (2, 512, 8, 8, 1, 4, (2, 2, 1, 1), 1, False),       # patch discriminator output
         (2, 64, 16, 16, 3, 7, (3, 3, 3, 3), 1, True),       # cyclegan generator output layer
         (2, 64, 20, 12, 1, 3, (1, 1, 1, 1), 1, False),      # dcgan generator output layer (un-fused form)
         (1, 32, 10, 10, 2, 3, (1, 1, 1, 1), 2, False),      # 3x3 behind the x2 upsample: folded taps, odd padding parity
         (2, 64, 8, 12, 3, 4, (1, 1, 2, 2), 2, False),       # 4x4 folded, odd padding parity
         (2, 128, 7, 5, 1, 4, (2, 2, 1, 1), 2, False),       # folded, ragged, one output channel
         (2, 256, 6, 6, 4, 3, (1, 1, 1, 1), 1, False)]


def _virtual_input(x, pads, up, reflect):
    """The tensor the convolution slides over: upsampled, then padded (top, left, bottom, right)."""
    if up == 2:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    t, l, b, r = pads
    return F.pad(x, (l, r, t, b), mode="reflect" if reflect else "constant")


@pytest.mark.parametrize("case", CASES)
def test_fewk_forward_and_gradients(case):
    from b200gan import ops
    from b200gan.functional import ACT_TANH, ALGO_SIMT, PACK_SIMT_DGRAD, PACK_SIMT_FPROP
    n, c, h, w_, k, ks, pads, up, reflect = case
    torch.manual_seed(5)
    x = torch.randn(n, c, h, w_, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(k, c, ks, ks, device="cuda") * 0.05).requires_grad_(True)
    b = (torch.randn(k, device="cuda") * 0.1).requires_grad_(True)
    g, _ = ops.make_geom((n, c, h, w_), (k, c, ks, ks), 1, pads, 1 if reflect else 0, up)
    ref = torch.tanh(F.conv2d(_virtual_input(x, pads, up, reflect), w, b))
    assert tuple(ref.shape) == (n, k, g.P, g.Q)
    y = ops.conv_fprop(g, x.detach(), ops.pack_weights(g, w.detach(), PACK_SIMT_FPROP), ALGO_SIMT, bias=b.detach(),
                       act=ACT_TANH)
    assert rel_err(y, ref) < TOL
    # gradients of the pre-activation z: dz -> (dx, dw)
    dz = torch.randn_like(ref).contiguous(memory_format=torch.channels_last)
    z = F.conv2d(_virtual_input(x, pads, up, reflect), w, b)
    dx_ref, dw_ref = torch.autograd.grad(z, (x, w), dz)
    dw, db = ops.conv_wgrad(g, x.detach(), dz, tuple(w.shape), True, ALGO_SIMT)
    assert rel_err(dw, dw_ref) < TOL
    assert rel_err(db, dz.sum((0, 2, 3))) < TOL
    dx = ops.conv_dgrad(g, dz, ops.pack_weights(g, w.detach(), PACK_SIMT_DGRAD), ALGO_SIMT)
    assert rel_err(dx, dx_ref) < TOL


# (N, C, H, W, K, kernel, stride, pad, reflect): layers with a handful of INPUT channels -- the staged SIMT kernels of
# csrc/narrow_block.cu on their own (nb_plain_fprop / nb_plain_dgrad)
EDGE_CASES = [(2, 3, 32, 32, 64, 4, 2, 1, False),      # pix2pix generator down1 (models.py:76)
              (2, 6, 32, 32, 64, 4, 2, 1, False),      # pix2pix discriminator block 1 (models.py:118)
              (3, 6, 20, 28, 64, 4, 2, 1, False),      # ragged
              (2, 3, 24, 24, 64, 7, 1, 3, True),       # cyclegan stem (models.py:49-50)
              (2, 3, 16, 16, 32, 3, 2, 1, False),
              (2, 4, 16, 16, 16, 3, 1, 1, False)]


@pytest.mark.parametrize("case", EDGE_CASES)
def test_few_input_channel_layers(case):
    from b200gan import ops
    from b200gan.functional import ACT_LRELU, ALGO_SIMT, PACK_SIMT_DGRAD, PACK_SIMT_FPROP
    n, c, h, w_, k, ks, st, pad, reflect = case
    torch.manual_seed(6)
    x = torch.randn(n, c, h, w_, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(k, c, ks, ks, device="cuda") * 0.1).requires_grad_(True)
    b = torch.randn(k, device="cuda") * 0.1
    g, _ = ops.make_geom((n, c, h, w_), (k, c, ks, ks), st, (pad,) * 4, 1 if reflect else 0)
    xv = F.pad(x, (pad,) * 4, mode="reflect") if reflect else x
    z = F.conv2d(xv, w, b, st, 0 if reflect else pad)
    y = ops.conv_fprop(g, x.detach(), ops.pack_weights(g, w.detach(), PACK_SIMT_FPROP), ALGO_SIMT, bias=b, act=ACT_LRELU,
                       slope=0.2)
    assert rel_err(y, F.leaky_relu(z, 0.2)) < TOL
    dz = torch.randn_like(z).contiguous(memory_format=torch.channels_last)
    dx_ref, dw_ref = torch.autograd.grad(z, (x, w), dz)
    dx = ops.conv_dgrad(g, dz, ops.pack_weights(g, w.detach(), PACK_SIMT_DGRAD), ALGO_SIMT)
    assert rel_err(dx, dx_ref) < TOL
    dw, _ = ops.conv_wgrad(g, x.detach(), dz, tuple(w.shape), False, ALGO_SIMT)
    assert rel_err(dw, dw_ref) < TOL
