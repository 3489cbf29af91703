"""Bring-up probe for the tcgen05 conv kernel: runs one geometry through the TC path and the SIMT path
(both via the C ABI) plus stock torch, prints error statistics.  Each case runs in its own process under a
timeout so that a hung kernel cannot take the whole session down.

    python tools/tc_probe.py            # all cases
    python tools/tc_probe.py CASE_JSON  # one case (internal)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-gan_b200"))

CASES = [
    dict(name="DCGAN D conv4 s2 k3 64->128 8x8 n16 (16-pixel grid)", cin=64, cout=128, k=3, pad=1, h=8, w=8, n=16, up=1, stride=2),
    dict(name="convT k4 s2 512->64 2x2 n2 (4-pixel grid)", cin=512, cout=64, k=4, pad=1, h=2, w=2, n=2, up=1, stride=2, transposed=True),
    dict(name="pix2pix down8 k4 s2 512->512 2x2 n3", cin=512, cout=512, k=4, pad=1, h=2, w=2, n=3, up=1, stride=2),
    dict(name="s2 k3 64->128 16x16 n4", cin=64, cout=128, k=3, pad=1, h=16, w=16, n=4, up=1, stride=2),
    dict(name="s2 k4 128->256 32x32 n2", cin=128, cout=256, k=4, pad=1, h=32, w=32, n=2, up=1, stride=2),
    dict(name="s2 k3 32->64 16x16 n8 (BN=32 dgrad)", cin=32, cout=64, k=3, pad=1, h=16, w=16, n=8, up=1, stride=2),
    dict(name="convT k4 s2 256->128 8x8 n2", cin=256, cout=128, k=4, pad=1, h=8, w=8, n=2, up=1, stride=2, transposed=True),
    dict(name="convT k4 s2 64->32 8x8 n3", cin=64, cout=32, k=4, pad=1, h=8, w=8, n=3, up=1, stride=2, transposed=True),
    dict(name="fprop 64->64 3x3 16x16 n2", cin=64, cout=64, k=3, pad=1, h=16, w=16, n=2, up=1),
    dict(name="fprop 64->128 3x3 8x8 n3", cin=64, cout=128, k=3, pad=1, h=8, w=8, n=3, up=1),
    dict(name="fprop 128->128 3x3 32x32 n2", cin=128, cout=128, k=3, pad=1, h=32, w=32, n=2, up=1),
    dict(name="fprop up2 128->128 16x16 n4", cin=128, cout=128, k=3, pad=1, h=16, w=16, n=4, up=2),
    dict(name="fprop up2 128->64 32x32 n2", cin=128, cout=64, k=3, pad=1, h=32, w=32, n=2, up=2),
    dict(name="fprop 128->64 3x3 24x20 n3 ragged", cin=128, cout=64, k=3, pad=1, h=24, w=20, n=3, up=1),
]


def run_case(c):
    import torch
    from b200gan import ops
    from b200gan._lib import (ALGO_SIMT, ALGO_TC, PACK_SIMT_DGRAD, PACK_SIMT_FPROP, PACK_TC_DGRAD, PACK_TC_DGRAD_UP2,
                              PACK_TC_FPROP, PACK_TC_FPROP_UP2)
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    n, cin, cout, k, pad, h, w, up = c["n"], c["cin"], c["cout"], c["k"], c["pad"], c["h"], c["w"], c["up"]
    stride, tr = c.get("stride", 1), c.get("transposed", False)
    x = torch.randn(n, cin, h, w, device="cuda").contiguous(memory_format=torch.channels_last)
    wt = torch.randn(*((cin, cout, k, k) if tr else (cout, cin, k, k)), device="cuda") * 0.05
    g, oshape = ops.make_geom(tuple(x.shape), tuple(wt.shape), stride, (pad, pad, pad, pad), 0, up, tr)
    F = torch.nn.functional
    conv = (lambda a, b: F.conv_transpose2d(a, b, None, stride, pad)) if tr else (lambda a, b: F.conv2d(a, b, None, stride, pad))
    print("  tc_supported fprop/dgrad:", ops.tc_supported(g, 0), ops.tc_supported(g, 1), "out", oshape, flush=True)
    xin = torch.nn.functional.interpolate(x, scale_factor=2) if up == 2 else x
    ref = conv(xin, wt)
    y_s = ops.conv_fprop(g, x, ops.pack_weights(g, wt, PACK_SIMT_FPROP), ALGO_SIMT)
    torch.cuda.synchronize()

    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm()).item()

    print(f"  SIMT fprop rel err vs torch: {rel(y_s, ref):.3e}", flush=True)
    y_t = ops.conv_fprop(g, x, ops.pack_weights(g, wt, PACK_TC_FPROP_UP2 if up == 2 else PACK_TC_FPROP), ALGO_TC)
    torch.cuda.synchronize()
    e = rel(y_t, ref)
    print(f"  TC   fprop rel err vs torch: {e:.3e}   max|y| {ref.abs().max().item():.3f} "
          f"max|diff| {(y_t - ref).abs().max().item():.3e} nan {bool(torch.isnan(y_t).any())}", flush=True)
    if e > 2e-3:
        d = (y_t - ref).abs()
        bad = (d > 1e-2 * ref.abs().max()).nonzero()
        print("  first mismatching (n,k,p,q):", bad[:8].tolist(), " count", bad.shape[0], "of", d.numel())
        print("  ours", y_t.flatten()[:6].tolist(), "\n  ref ", ref.flatten()[:6].tolist())
    dy = torch.randn_like(ref).contiguous(memory_format=torch.channels_last)
    xin_r = xin.clone().requires_grad_(True)
    conv(xin_r, wt).backward(dy)
    dref = xin_r.grad
    if up == 2:
        dref = dref.view(n, cin, h, 2, w, 2).sum(dim=(3, 5))
    dx_s = ops.conv_dgrad(g, dy, ops.pack_weights(g, wt, PACK_SIMT_DGRAD), ALGO_SIMT)
    print(f"  SIMT dgrad rel err: {rel(dx_s, dref):.3e}", flush=True)
    dx_t = ops.conv_dgrad(g, dy, ops.pack_weights(g, wt, PACK_TC_DGRAD_UP2 if up == 2 else PACK_TC_DGRAD), ALGO_TC)
    torch.cuda.synchronize()
    print(f"  TC   dgrad rel err: {rel(dx_t, dref):.3e}", flush=True)
    w_r = wt.clone().requires_grad_(True)
    conv(xin, w_r).backward(dy)
    dw_s, _ = ops.conv_wgrad(g, x, dy, tuple(wt.shape), False, ALGO_SIMT)
    print(f"  SIMT wgrad rel err: {rel(dw_s, w_r.grad):.3e}", flush=True)
    if ops.tc_supported(g, 2):
        for variant in ("0", "1"):
            os.environ["B200GAN_WG_VARIANT"] = variant
            dw_t, _ = ops.conv_wgrad(g, x, dy, tuple(wt.shape), False, ALGO_TC)
            torch.cuda.synchronize()
            print(f"  TC   wgrad variant {variant} rel err: {rel(dw_t, w_r.grad):.3e}", flush=True)
        os.environ.pop("B200GAN_WG_VARIANT")
    else:
        print("  TC wgrad not supported for this geometry")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(json.loads(sys.argv[1]))
        sys.exit(0)
    for c in CASES:
        print("==", c["name"], flush=True)
        try:
            r = subprocess.run([sys.executable, __file__, json.dumps(c)], timeout=60, capture_output=True, text=True)
            print(r.stdout, end="")
            if r.returncode != 0:
                print("  EXIT", r.returncode, r.stderr[-1500:])
        except subprocess.TimeoutExpired as e:
            print("  TIMEOUT (kernel hang?)", (e.stdout or b"")[-500:] if e.stdout else "")
