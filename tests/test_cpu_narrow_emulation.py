"""CPU emulation of the experimental narrow-layer convolution kernels (pytorch-gan_b200/csrc/conv_narrow.cuh).

The per-thread bodies of those CUDA kernels are plain C++; tests/emu/narrow_emu.cpp compiles the SAME source with g++ and
runs every (parity class, block, thread) of a launch sequentially, including the out-of-range tail threads.  Here the
result is compared with torch on CPU for the geometries the kernels are meant for (DCGAN discriminator 1->16->32->64,
k3 s2 p1, dcgan.py:77-88) and a few others.  This pins the index arithmetic (parity classes of the strided transposed
gather, ragged sizes, epilogue order) without a GPU.  The kernels stay behind B200GAN_NARROW=1 until they have been run
and timed on hardware.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "narrow_emu.cpp")
HDR = os.path.join(HERE, "..", "pytorch-gan_b200", "csrc", "conv_narrow.cuh")
OUT = os.path.join(HERE, "emu", "_build", "libnarrow_emu.so")

FP = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def emu():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    newest = max(os.path.getmtime(SRC), os.path.getmtime(HDR))
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < newest:
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", OUT, SRC], check=True)
    lib = ctypes.CDLL(OUT)
    lib.emu_narrow_gather.restype = ctypes.c_int
    lib.emu_narrow_gather.argtypes = [ctypes.c_int] * 13 + [FP, FP, ctypes.c_int, ctypes.c_float, ctypes.c_int, FP, FP, FP,
                                                            ctypes.c_int, ctypes.c_int]
    lib.emu_narrow_wgrad.restype = ctypes.c_int
    lib.emu_narrow_wgrad.argtypes = [ctypes.c_int] * 12 + [FP, FP, FP, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return lib


def _aligned(t):
    """contiguous fp32 numpy copy on a 64-byte boundary (the vector paths of the kernels need 16)"""
    a = np.ascontiguousarray(t.detach().numpy().astype(np.float32))
    buf = np.empty(a.size + 16, dtype=np.float32)
    off = (-buf.ctypes.data % 64) // 4
    out = buf[off:off + a.size].reshape(a.shape)
    out[...] = a
    return out


def _ptr(a):
    return a.ctypes.data_as(FP) if a is not None else None


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _gather(emu, x_nhwc, wp, out_hw, K, R, S, stride, pad, mode, kt, block=64, bias=None, chan_scale=None, act=0,
            slope=0.0, rtf=0):
    n, h, w, c = x_nhwc.shape
    p, q = out_hw
    xa, wa = _aligned(x_nhwc), _aligned(wp)
    ba = _aligned(bias) if bias is not None else None
    ca = _aligned(chan_scale) if chan_scale is not None else None
    y = _aligned(torch.full((n, p, q, K), float("nan")))  # every element must be written exactly by the kernel
    rc = emu.emu_narrow_gather(n, h, w, c, p, q, K, R, S, stride, pad, pad, mode, _ptr(ba), _ptr(ca), act, slope, rtf,
                               _ptr(xa), _ptr(wa), _ptr(y), kt, block)
    assert rc == 0
    return torch.from_numpy(np.array(y)).permute(0, 3, 1, 2)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


CONV_CASES = [  # cin, cout, k, stride, pad, H, W, N
    (1, 16, 3, 2, 1, 12, 10, 2),    # D conv1 (scalar contraction path)
    (16, 32, 3, 2, 1, 10, 8, 2),    # D conv2
    (32, 64, 3, 2, 1, 6, 6, 3),     # D conv3
    (16, 32, 3, 2, 1, 11, 9, 2),    # odd maps: ragged parity classes in the dgrad
    (8, 16, 4, 2, 1, 8, 6, 2),      # k4 s2 p1 (pix2pix/models.py:23 geometry)
    (4, 8, 3, 1, 1, 7, 5, 2),       # stride 1
    (3, 8, 3, 1, 0, 6, 6, 1),       # odd channel count -> scalar path, no padding
    (16, 3, 3, 1, 1, 5, 5, 2),      # 3 outputs: KT = 1 only
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[str(c) for c in CONV_CASES])
def test_emulated_fprop_dgrad_wgrad_match_torch(emu, case):
    cin, cout, k, stride, pad, h, w, n = case
    torch.manual_seed(hash(case) % 1000)
    x = torch.randn(n, cin, h, w, requires_grad=True)
    wt = torch.randn(cout, cin, k, k, requires_grad=True)
    y = F.conv2d(x, wt, None, stride, pad)
    dy = torch.randn_like(y)
    gx, gw = torch.autograd.grad(y, (x, wt), dy)
    p, q = y.shape[2:]
    kts = [kt for kt in (8, 4, 1) if cout % kt == 0]
    # forward: mode 0, weights [tap][cin][cout]
    for kt in kts:
        yo = _gather(emu, _nhwc(x), wt.permute(2, 3, 1, 0).contiguous(), (p, q), cout, k, k, stride, pad, 0, kt)
        assert _rel(yo, y) < 1e-5, ("fprop", kt)
    # input gradient: mode 1 over dy, weights [tap][cout][cin]; stride 2 -> four parity classes
    for kt in [kt for kt in (8, 4, 1) if cin % kt == 0]:
        for block in (64, 96):
            go = _gather(emu, _nhwc(dy), wt.permute(2, 3, 0, 1).contiguous(), (h, w), cin, k, k, stride, pad, 1, kt, block)
            assert _rel(go, gx) < 1e-5, ("dgrad", kt, block)
    # weight gradient in OIHW, any split of the pixel range
    xa, da = _aligned(_nhwc(x)), _aligned(_nhwc(dy))
    for dt in [d for d in (4, 1) if cout % d == 0]:
        for splits in (1, 3, 7):
            dw = _aligned(torch.zeros(cout, cin, k, k))
            rc = emu.emu_narrow_wgrad(n, h, w, cin, p, q, cout, k, k, stride, pad, pad, _ptr(xa), _ptr(da), _ptr(dw), dt,
                                      splits, 64)
            assert rc == 0
            assert _rel(torch.from_numpy(np.array(dw)), gw) < 1e-5, ("wgrad", dt, splits)


@pytest.mark.parametrize("cin,cout,h,w", [(16, 8, 4, 3), (8, 16, 1, 1), (4, 4, 5, 6)])
def test_emulated_conv_transpose_forward_and_backward(emu, cin, cout, h, w):
    """ConvTranspose2d k4 s2 p1 (pix2pix/models.py:39): forward is the strided transposed gather (mode 1) with
    weights [tap][cin][cout]; its input gradient is the plain gather (mode 0) over dy with [tap][cout][cin];
    its weight gradient swaps the roles of the two tensors (gathered = dy, dense = x) and lands in IOHW."""
    torch.manual_seed(3)
    n, k, stride, pad = 2, 4, 2, 1
    x = torch.randn(n, cin, h, w, requires_grad=True)
    wt = torch.randn(cin, cout, k, k, requires_grad=True)
    y = F.conv_transpose2d(x, wt, None, stride, pad)
    dy = torch.randn_like(y)
    gx, gw = torch.autograd.grad(y, (x, wt), dy)
    oh, ow = y.shape[2:]
    kt = 4
    yo = _gather(emu, _nhwc(x), wt.permute(2, 3, 0, 1).contiguous(), (oh, ow), cout, k, k, stride, pad, 1, kt)
    assert _rel(yo, y) < 1e-5
    go = _gather(emu, _nhwc(dy), wt.permute(2, 3, 1, 0).contiguous(), (h, w), cin, k, k, stride, pad, 0, kt)
    assert _rel(go, gx) < 1e-5
    # dw[ci][co][r][s] = sum x[n][ih][iw][ci] * dy[n][ih*2-1+r][iw*2-1+s][co]: gathered = dy (mode 0 addressing from the
    # low-resolution pixel grid), dense = x; the kernel writes [dense channel][gathered channel][r][s] = [ci][co][r][s]
    dw = _aligned(torch.zeros(cin, cout, k, k))
    rc = emu.emu_narrow_wgrad(n, oh, ow, cout, h, w, cin, k, k, stride, pad, pad, _ptr(_aligned(_nhwc(dy))),
                              _ptr(_aligned(_nhwc(x))), _ptr(dw), 4, 2, 64)
    assert rc == 0
    assert _rel(torch.from_numpy(np.array(dw)), gw) < 1e-5


def test_emulated_epilogue_order_bias_act_scale_round(emu):
    """bias -> activation -> Dropout2d channel scale -> TF32 rounding, the order of every other conv kernel
    (dcgan.py:77-80: Conv -> LeakyReLU -> Dropout2d)."""
    torch.manual_seed(5)
    n, cin, cout, h, w = 2, 16, 32, 8, 8
    x, wt, b = torch.randn(n, cin, h, w), torch.randn(cout, cin, 3, 3) * 0.1, torch.randn(cout)
    scale = (torch.rand(n, cout) > 0.25).float() / 0.75
    ref = F.leaky_relu(F.conv2d(x, wt, b, 2, 1), 0.2) * scale[:, :, None, None]
    yo = _gather(emu, _nhwc(x), wt.permute(2, 3, 1, 0).contiguous(), ref.shape[2:], cout, 3, 3, 2, 1, 0, 8, bias=b,
                 chan_scale=scale, act=1, slope=0.2)
    assert _rel(yo, ref) < 1e-5
    yr = _gather(emu, _nhwc(x), wt.permute(2, 3, 1, 0).contiguous(), ref.shape[2:], cout, 3, 3, 2, 1, 0, 8, bias=b,
                 chan_scale=scale, act=1, slope=0.2, rtf=1)
    bits = yr.contiguous().view(torch.int32)
    assert torch.all((bits & 0x1FFF) == 0)  # 13 low mantissa bits cleared
    assert _rel(yr, ref) < 6e-4               # and within half a TF32 ulp (2^-11) of the fp32 result
    for act, fn in ((2, torch.relu), (3, torch.tanh), (4, torch.sigmoid)):
        yo = _gather(emu, _nhwc(x), wt.permute(2, 3, 1, 0).contiguous(), ref.shape[2:], cout, 3, 3, 2, 1, 0, 4, bias=b,
                     act=act)
        assert _rel(yo, fn(F.conv2d(x, wt, b, 2, 1))) < 1e-5
