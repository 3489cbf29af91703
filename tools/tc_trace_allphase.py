"""Per-CTA clock64 timeline of conv_tc_up2_allphase_kernel (bring-up; same slot layout as tools/tc_trace.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-gan_b200"))
import torch  # noqa: E402

from b200gan import ops  # noqa: E402
from b200gan._lib import ALGO_TC, PACK_TC_FPROP_UP2  # noqa: E402


def run(cin, cout, h, w, n):
    x = torch.randn(n, cin, h, w, device="cuda").contiguous(memory_format=torch.channels_last)
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.02
    g, _ = ops.make_geom(tuple(x.shape), tuple(wt.shape), 1, (1, 1, 1, 1), 0, 2, False)
    packed = ops.pack_weights(g, wt, PACK_TC_FPROP_UP2)
    for _ in range(3):
        ops.conv_fprop(g, x, packed, ALGO_TC)
    torch.cuda.synchronize()
    ncta = 4096
    trace = torch.zeros(ncta * 64, device="cuda", dtype=torch.int64)
    os.environ["B200GAN_TC_TRACE"] = str(trace.data_ptr())
    ops.conv_fprop(g, x, packed, ALGO_TC)
    torch.cuda.synchronize()
    os.environ.pop("B200GAN_TC_TRACE")
    t = trace.view(ncta, 64).cpu()
    used = (t[:, 0] != 0).nonzero().flatten()
    t0 = t[used, 0].min().item()
    f = lambda a, b: (t[used, a] - t[used, b]).float()  # noqa: E731
    print(f"== fprop up2 all-phase {cin}->{cout} low-res {h}x{w} n{n}: {len(used)} CTAs; kernel span "
          f"{t[used, 42].max().item() - t0} cycles")
    for name, v in [("lifetime", f(42, 0)), ("setup", f(1, 0)), ("main loop (setup -> accumulators complete)", f(40, 1)),
                    ("first 15 producer intervals /15", f(17, 2) / 15), ("first 15 mma-full intervals /15", f(35, 20) / 15),
                    ("first load latency (issue -> full)", f(20, 2)), ("last issue -> last full", f(36, 18)),
                    ("last full -> epilogue start", f(40, 36)), ("epilogue (start -> stores read)", f(41, 40)),
                    ("  phase 0: ld+math+stage", f(49, 48)), ("  phase 0: fence+bar", f(50, 49)),
                    ("  phase 1 total", f(51, 48)), ("  phase 3 total", f(43, 57)), ("  final store drain", f(41, 43)),
                    ("epilogue end -> exit", f(42, 41))]:
        print(f"   {name:46s} mean {v.mean():8.0f}  min {v.min():8.0f}  max {v.max():8.0f}")
    for cta in [used[0].item(), used[len(used) // 2].item(), used[-1].item()]:
        r = t[cta]
        s0 = r[0].item()
        print(f"   CTA {cta}: start +{s0 - t0}")
        print("      producer issue:", [(r[2 + i].item() - s0) for i in range(16)], "last", r[18].item() - s0)
        print("      mma got full  :", [(r[20 + i].item() - s0) for i in range(16)], "last", r[36].item() - s0)
        print("      epilogue      :", {k: r[k].item() - s0 for k in (40, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 43, 41, 42)})


if __name__ == "__main__":
    run(128, 64, 32, 32, 128)
