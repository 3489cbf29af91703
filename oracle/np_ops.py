"""numpy restatement of the torch operators on the hot path, written
from their mathematical definitions with plain loops / einsum -- independent of torch so that it
can referee both torch-CPU and the CUDA kernels on small cases.  TEST INFRASTRUCTURE ONLY.
All tensors are NCHW float64 numpy arrays.
"""
import numpy as np


def pad2d(x, pads, mode="zero"):
    """pads = (top, left, bottom, right).  zero: nn.ZeroPad2d (pix2pix/models.py:78);
    reflect: nn.ReflectionPad2d (cyclegan/models.py:27), edge pixel not repeated."""
    t, l, b, r = pads
    if mode == "zero":
        return np.pad(x, ((0, 0), (0, 0), (t, b), (l, r)))
    return np.pad(x, ((0, 0), (0, 0), (t, b), (l, r)), mode="reflect")


def upsample2x(x):
    """nn.Upsample(scale_factor=2), nearest: dst -> src = dst // 2 (dcgan.py:54)."""
    return x.repeat(2, axis=2).repeat(2, axis=3)


def conv2d(x, w, b=None, stride=1, pad=0):
    """nn.Conv2d (dcgan.py:55,78): cross-correlation, w [K,C,R,S]."""
    x = pad2d(x, (pad, pad, pad, pad))
    n, c, h, wd = x.shape
    k, _, r, s = w.shape
    p = (h - r) // stride + 1
    q = (wd - s) // stride + 1
    y = np.zeros((n, k, p, q))
    for i in range(r):
        for j in range(s):
            patch = x[:, :, i:i + (p - 1) * stride + 1:stride, j:j + (q - 1) * stride + 1:stride]
            y += np.einsum("nchw,kc->nkhw", patch, w[:, :, i, j])
    if b is not None:
        y += b[None, :, None, None]
    return y


def conv_transpose2d(x, w, b=None, stride=2, pad=1):
    """nn.ConvTranspose2d (pix2pix/models.py:39): w [C,K,R,S]; out = (H-1)*stride - 2*pad + R."""
    n, c, h, wd = x.shape
    _, k, r, s = w.shape
    full = np.zeros((n, k, (h - 1) * stride + r, (wd - 1) * stride + s))
    for i in range(r):
        for j in range(s):
            full[:, :, i:i + (h - 1) * stride + 1:stride, j:j + (wd - 1) * stride + 1:stride] += np.einsum(
                "nchw,ck->nkhw", x, w[:, :, i, j])
    hh, ww = full.shape[2], full.shape[3]
    y = full[:, :, pad:hh - pad, pad:ww - pad]
    if b is not None:
        y = y + b[None, :, None, None]
    return y


def batch_norm_train(x, gamma, beta, eps):
    """nn.BatchNorm2d in training mode (dcgan.py:56): biased variance for normalisation.
    Returns (y, batch_mean, unbiased_var) -- the latter two feed the running statistics."""
    mean = x.mean(axis=(0, 2, 3))
    var = x.var(axis=(0, 2, 3))
    cnt = x.shape[0] * x.shape[2] * x.shape[3]
    y = (x - mean[None, :, None, None]) / np.sqrt(var[None, :, None, None] + eps)
    y = y * gamma[None, :, None, None] + beta[None, :, None, None]
    return y, mean, var * cnt / max(cnt - 1, 1)


def instance_norm(x, eps=1e-5):
    """nn.InstanceNorm2d(C) (pix2pix/models.py:25): affine=False, per (n, c) statistics."""
    mean = x.mean(axis=(2, 3), keepdims=True)
    var = x.var(axis=(2, 3), keepdims=True)
    return (x - mean) / np.sqrt(var + eps)


def leaky_relu(x, slope=0.2):
    return np.where(x > 0, x, x * slope)


def gp_mlp_closed_form(xi, w1, b1, w2, b2, w3, slope=0.2, lam=10.0):
    """Closed form of wgan_gp.py:119-138 + its double backward for the critic of wgan_gp.py:72-78
    (LeakyReLU'' = 0 a.e.).  float64 numpy.  Returns (lam*gp, dW1, dW2, dW3)."""
    n = xi.shape[0]
    x = xi.reshape(n, -1)
    h1 = x @ w1.T + b1
    m1 = np.where(h1 > 0, 1.0, slope)
    h2 = (h1 * m1) @ w2.T + b2
    m2 = np.where(h2 > 0, 1.0, slope)
    g2 = w3.reshape(1, -1) * m2
    g1 = (g2 @ w2) * m1
    gx = g1 @ w1
    r = np.sqrt((gx * gx).sum(1))
    gp = lam * np.mean((r - 1.0) ** 2)
    u = (lam * 2.0 / n) * ((r - 1.0) / r)[:, None] * gx
    dw1 = g1.T @ u
    t = (u @ w1.T) * m1
    dw2 = g2.T @ t
    dw3 = ((t @ w2.T) * m2).sum(0)
    return gp, dw1, dw2, dw3
