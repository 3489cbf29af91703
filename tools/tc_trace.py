"""Per-CTA clock64 timeline of conv_tc_kernel (bring-up).  Sets B200GAN_TC_TRACE to a device buffer address."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-gan_b200"))
import torch  # noqa: E402

from b200gan import ops  # noqa: E402
from b200gan._lib import ALGO_TC, PACK_TC_FPROP_UP2, PACK_TC_DGRAD_UP2  # noqa: E402


def run(cin, cout, h, w, n, dgrad=False):
    x = torch.randn(n, cin, h, w, device="cuda").contiguous(memory_format=torch.channels_last)
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.02
    g, oshape = ops.make_geom(tuple(x.shape), tuple(wt.shape), 1, (1, 1, 1, 1), 0, 2, False)
    dy = torch.randn(oshape, device="cuda").contiguous(memory_format=torch.channels_last)
    packed = ops.pack_weights(g, wt, PACK_TC_DGRAD_UP2 if dgrad else PACK_TC_FPROP_UP2)
    fn = (lambda: ops.conv_dgrad(g, dy, packed, ALGO_TC)) if dgrad else (lambda: ops.conv_fprop(g, x, packed, ALGO_TC))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ncta = 8192
    trace = torch.zeros(ncta * 64, device="cuda", dtype=torch.int64)
    os.environ["B200GAN_TC_TRACE"] = str(trace.data_ptr())
    fn()
    torch.cuda.synchronize()
    os.environ.pop("B200GAN_TC_TRACE")
    t = trace.view(ncta, 64).cpu()
    used = (t[:, 0] != 0).nonzero().flatten()
    t0 = t[used, 0].min().item()
    print(f"== {'dgrad' if dgrad else 'fprop'} up2 {cin}->{cout} low-res {h}x{w} n{n}: {len(used)} CTAs traced; cycles relative to first CTA start")
    life = (t[used, 42] - t[used, 0]).float()
    print(f"   CTA lifetime cycles: mean {life.mean():.0f} min {life.min():.0f} max {life.max():.0f}; kernel span {(t[used, 42].max().item() - t0)} cycles")
    for cta in [used[0].item(), used[len(used) // 2].item(), used[-1].item()]:
        r = t[cta]
        s0 = r[0].item()
        prod = [(r[2 + i].item() - s0) for i in range(16) if r[2 + i].item()]
        mma = [(r[20 + i].item() - s0) for i in range(16) if r[20 + i].item()]
        print(f"   CTA {cta}: start +{s0 - t0}  setup {r[1].item() - s0}  epi_start {r[40].item() - s0}  epi_end {r[41].item() - s0}  exit {r[42].item() - s0}")
        epi = {k: (r[k].item() - s0) for k in (43, 44, 45) if r[k].item()}
        chunks = [(r[48 + 3 * i].item() - s0, r[49 + 3 * i].item() - s0, r[50 + 3 * i].item() - s0) for i in range(4) if r[48 + 3 * i].item()]
        print("      epilogue: loop_done/bar_done/tma_issued", epi, " per chunk (before ld, after ld, after st.shared):", chunks)
        print("      producer issue:", prod)
        print("      mma got full  :", mma)


if __name__ == "__main__":
    run(128, 128, 16, 16, 128)
    run(128, 64, 32, 32, 128)
    run(128, 128, 16, 16, 128, dgrad=True)
