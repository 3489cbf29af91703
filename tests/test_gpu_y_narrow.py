"""GPU parity of the EXPERIMENTAL narrow-layer kernels (csrc/conv_narrow.cu).  They are dispatched only when the process
starts with B200GAN_NARROW=1, so this module is skipped in a default run; `B200GAN_NARROW=1 pytest -m gpu` runs it AND
sends every <= 64-channel SIMT convolution of the whole GPU suite through them.  Their index arithmetic is already
pinned on CPU (tests/test_cpu_narrow_emulation.py)."""
import os

import pytest
import torch

from conftest import rel_err

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("B200GAN_NARROW", "0") in ("", "0"),
                                 reason="narrow-layer kernels are opt-in (B200GAN_NARROW=1) until validated on hardware")]

CASES = [  # cin, cout, k, stride, pad, H, W, N
    (1, 16, 3, 2, 1, 64, 64, 8),     # dcgan D conv1
    (16, 32, 3, 2, 1, 32, 32, 8),    # D conv2
    (32, 64, 3, 2, 1, 16, 16, 8),    # D conv3
    (1, 64, 3, 1, 1, 32, 32, 4),     # one input channel, stride 1 (the dgrad direction of the G output conv)
    (16, 32, 3, 2, 1, 11, 9, 3),     # ragged parity classes
    (8, 16, 4, 2, 1, 12, 10, 2),
]


@pytest.fixture(autouse=True)
def _fp32_simt():
    from b200gan import ops
    old = ops.Config.algo
    ops.Config.algo = "simt"  # keep the tensor-core path out: everything below goes SIMT -> narrow
    torch.backends.cudnn.allow_tf32 = False
    yield
    ops.Config.algo = old


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_narrow_conv_fwd_bwd(case):
    from b200gan import nn as bnn
    cin, cout, k, stride, pad, h, w, n = case
    torch.manual_seed(1)
    ref = torch.nn.Sequential(torch.nn.Conv2d(cin, cout, k, stride, pad), torch.nn.LeakyReLU(0.2)).cuda()
    ours = bnn.Sequential(bnn.Conv2d(cin, cout, k, stride, pad), bnn.LeakyReLU(0.2)).cuda()
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(n, cin, h, w, device="cuda")
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yr, yo = ref(xr), ours(xo)
    assert rel_err(yo, yr) < 1e-4
    gy = torch.randn_like(yr)
    yr.backward(gy)
    yo.backward(gy)
    assert rel_err(xo.grad, xr.grad) < 1e-4
    assert rel_err(ours[0].weight.grad, ref[0].weight.grad) < 1e-4
    assert rel_err(ours[0].bias.grad, ref[0].bias.grad) < 1e-4


def test_narrow_conv_transpose_fwd_bwd():
    from b200gan import nn as bnn
    torch.manual_seed(2)
    ref = torch.nn.ConvTranspose2d(16, 8, 4, 2, 1, bias=False).cuda()
    ours = bnn.ConvTranspose2d(16, 8, 4, 2, 1, bias=False).cuda()
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(3, 16, 7, 5, device="cuda")
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yr, yo = ref(xr), ours(xo)
    assert rel_err(yo, yr) < 1e-4
    gy = torch.randn_like(yr)
    yr.backward(gy)
    yo.backward(gy)
    assert rel_err(xo.grad, xr.grad) < 1e-4
    assert rel_err(ours.weight.grad, ref.weight.grad) < 1e-4
