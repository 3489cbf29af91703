#!/usr/bin/env python
"""bench.py -- DCGAN 64x64, batch 128 per GPU, full G+D training step (dcgan.py:146-183) in images/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL).  Rank 0 prints ONE JSON line.  See the contract
in DESIGN.md section "Measurement".  Workload = BASELINE.json configs[1]; weak scaling (bs 128 per GPU).

  value        step loop replayed from a CUDA graph, inputs already resident in HBM (a pool of distinct
               batches; the per-step working set of ~1.5 GB of activations exceeds the 126 MB L2)
  e2e          the same step driven from pinned HOST buffers: H2D copy of images + z every step and a
               D2H read of the two losses (the reference's own per-step .item(), dcgan.py:185-188)
  roofline     the tcgen05 implicit-GEMM conv kernel of G conv2 timed alone with CUDA events
  cpu_baseline the oracle restatement of the reference step on the host cores (stock torch CPU)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "pytorch-gan_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

IMG, BATCH, LATENT = 64, 128, 100
# useful conv/linear FLOPs of one DCGAN step at bs 128 (reference formulation)
GFLOP_PER_STEP_REFERENCE_FORM = 359.0
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel (conv_tc_up2_allphase_kernel) on this
# workload, from the committed `ncu --set full` capture profiles/r1_ncu_full_up2_allphase_kernel.csv: 68.1 MB read +
# 78.4 MB written (algorithmic: 67 MB in + 134 MB out + 0.5 MB weights; part of the output is still in the 126 MB L2
# when the kernel ends; the per-phase predecessor conv_tc_kernel<64,4> moved 214 + 104 MB)
NCU_DRAM_TRAFFIC_BYTES_PER_LAUNCH = 146.5e6


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "stock", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager step loop (debugging)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.path = tempfile.mktemp(prefix="b200gan_clocks_", suffix=".csv")
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower() == "active":
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
def cpu_step_throughput(steps, warmup, threads=None, budget_s=150.0):
    """The reference's own CPU path: oracle restatement of dcgan.py:146-183 with stock torch.nn.
    Bounded sample: if `steps` full-batch steps would not fit in `budget_s`, every step processes a smaller batch
    of the same workload (per-image CPU cost is batch-insensitive); images/s is reported either way."""
    import torch
    from oracle import ref_models
    threads = threads or min(os.cpu_count() or 1, 32)   # more threads than that only add contention on this path
    torch.set_num_threads(threads)
    g, d = ref_models.build_dcgan(IMG, seed=0)
    og, od = ref_models.make_adam(g.parameters()), ref_models.make_adam(d.parameters())
    batch = BATCH
    imgs = ref_models.synthetic_images(batch, 1, IMG, IMG, seed=0)
    z = ref_models.synthetic_z(batch, seed=0)
    t0 = time.perf_counter()
    ref_models.dcgan_step(g, d, og, od, imgs, z)          # first warm-up step doubles as the probe
    t_probe = time.perf_counter() - t0
    if (steps + max(warmup - 1, 0)) * t_probe > budget_s:
        batch = int(BATCH * budget_s / ((steps + max(warmup - 1, 0)) * t_probe)) // 8 * 8
        batch = max(8, min(BATCH, batch))
        imgs, z = imgs[:batch], z[:batch]
    for _ in range(max(warmup - 1, 0)):
        ref_models.dcgan_step(g, d, og, od, imgs, z)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref_models.dcgan_step(g, d, og, od, imgs, z)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, threads, batch


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ips, spstep, threads, batch = cpu_step_throughput(args.steps, max(args.warmup, 1))
    sample = (f"{args.steps} steps of dcgan.py:146-183 on {batch} of the {BATCH} images per step ({IMG}x{IMG}) after "
              f"{args.warmup} warm-up, stock torch CPU, {threads} threads")
    print(json.dumps({
        "impl": "reference", "metric": "DCGAN 64x64 images/sec (full G+D step)", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": spstep * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DCGAN 64x64 synthetic, batch 128, reference CPU path (oracle port)",
                   "global_batch": BATCH},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------------
def _event_time(torch, fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def time_conv_kernel(torch, iters=20):
    """Roofline leg: the tcgen05 kernels alone on G conv2 of the step (Upsample + Conv 128->64 on [128,128,32,32],
    dcgan.py:58-59) as the folded 4-phase implicit GEMM.  Executed FLOPs (after the 2.25x upsample fold) /
    CUDA-event time.  In + out = 67 + 134 MB > L2, so no flush is needed between launches."""
    from b200gan import ops
    from b200gan._lib import ALGO_TC, PACK_TC_DGRAD_UP2, PACK_TC_FPROP_UP2
    x = torch.randn(BATCH, 128, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last)
    w = torch.randn(64, 128, 3, 3, device="cuda") * 0.02
    g, oshape = ops.make_geom(tuple(x.shape), tuple(w.shape), 1, (1, 1, 1, 1), 0, 2, False)
    if not ops.tc_supported(g, 0):
        return None
    dy = torch.randn(oshape, device="cuda").contiguous(memory_format=torch.channels_last)
    pf, pd = ops.pack_weights(g, w, PACK_TC_FPROP_UP2), ops.pack_weights(g, w, PACK_TC_DGRAD_UP2)
    flops_exec = 2.0 * BATCH * 32 * 32 * 4 * 64 * 128 * 4  # 4 phases x 4 taps x Cin 128 x Cout 64 per low-res pixel
    bytes_alg = (x.numel() + dy.numel() + pf.numel()) * 4
    ms_f = _event_time(torch, lambda: ops.conv_fprop(g, x, pf, ALGO_TC), iters)
    ms_d = _event_time(torch, lambda: ops.conv_dgrad(g, dy, pd, ALGO_TC), iters)
    ms_w = _event_time(torch, lambda: ops.conv_wgrad(g, x, dy, tuple(w.shape), False, ALGO_TC), iters)
    # context: cuBLAS TF32 GEMM on the same box (what "TF32 tensor peak" means in practice here)
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device="cuda")
    ms_g = _event_time(torch, lambda: a @ a, 5)
    return {"ms": ms_f, "tflops": flops_exec / ms_f / 1e9, "gbs": bytes_alg / ms_f / 1e6, "flops": flops_exec,
            "bytes": bytes_alg, "dgrad_tflops": flops_exec / ms_d / 1e9, "wgrad_tflops": flops_exec / ms_w / 1e9,
            "cublas_tf32_gemm_tflops": 2 * 8192.0 ** 3 / ms_g / 1e9}


def _finish(world):
    """Leave without tearing NCCL down: destroy_process_group() can block while CUDA graphs that captured
    collectives are still alive, and the measurement is already printed."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import b200gan
    from b200gan import _lib, train, zoo

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    b200gan.load_library()

    torch.manual_seed(0)  # identical init on every rank
    stock = args.impl == "stock"
    ns = zoo.namespace(stock=stock)
    if stock:  # the reference's own GPU path: stock torch.nn on cuDNN/cuBLAS with TF32 allowed (torch's conv default)
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.benchmark = True
    g = zoo.DCGANGenerator(IMG, nn=ns).to(dev)
    d = zoo.DCGANDiscriminator(IMG, nn=ns).to(dev)
    g.apply(zoo.weights_init_normal)
    d.apply(zoo.weights_init_normal)
    opt_g = torch.optim.Adam(g.parameters(), lr=2e-4, betas=(0.5, 0.999), capturable=True)
    opt_d = torch.optim.Adam(d.parameters(), lr=2e-4, betas=(0.5, 0.999), capturable=True)
    loss = torch.nn.BCELoss()
    valid = torch.ones(BATCH, 1, device=dev)
    fake = torch.zeros(BATCH, 1, device=dev)

    reduce_g = reduce_d = None
    if world > 1:
        from b200gan import ddp
        reduce_g = ddp.GradReducer(list(g.parameters()), world)
        reduce_d = ddp.GradReducer(list(d.parameters()), world)

    def step(imgs, z):
        gl, dl, _ = train.dcgan_step(g, d, opt_g, opt_d, imgs, z, loss, valid, fake, reduce_g, reduce_d)
        return torch.stack([gl, dl])

    # synthetic data: range of Normalize([.5],[.5]) (dcgan.py:126), z ~ N(0,1) (dcgan.py:160); per-rank streams
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    pool_n = 8
    host_imgs = [(torch.rand(BATCH, 1, IMG, IMG, generator=gen) * 2 - 1).pin_memory() for _ in range(pool_n)]
    host_z = [torch.randn(BATCH, LATENT, generator=gen).pin_memory() for _ in range(pool_n)]
    dev_imgs = [t.to(dev) for t in host_imgs]
    dev_z = [t.to(dev) for t in host_z]

    torch.manual_seed(99 + rank)  # Dropout2d streams differ per rank
    calls0 = _lib.CALLS
    if args.no_graph:
        runner = lambda a, b: step(a, b)  # noqa: E731
        for _ in range(3):
            step(dev_imgs[0], dev_z[0])
        calls_per_step = None
    else:
        runner = train.GraphedStep(step, [dev_imgs[0], dev_z[0]], warmup=3)
        calls_per_step = (_lib.CALLS - calls0) // 4  # 3 warm-up + 1 captured executions

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs ----
    for i in range(max(args.warmup, 3)):
        runner(dev_imgs[i % pool_n], dev_z[i % pool_n])
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = runner(dev_imgs[i % pool_n], dev_z[i % pool_n])
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    losses = out.tolist()

    # ---- e2e: host buffers in, losses out, every step ----
    for i in range(3):
        runner(host_imgs[i % pool_n], host_z[i % pool_n])
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(args.steps):
        out = runner(host_imgs[i % pool_n], host_z[i % pool_n])
        _ = out.cpu()  # D2H read of (g_loss, d_loss): the reference's per-step .item()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    if rank != 0:
        _finish(world)
        return

    hbm_peak, bf16_peak, peak_src = measured_peaks()
    tf32_peak = bf16_peak / 2.0  # kind::tf32 issues at half the kind::f16 rate (guide: 1.13 vs 2.25 PF nominal)
    conv = None if stock else time_conv_kernel(torch)
    roofline = None
    if conv is not None:
        roofline = {"kernel": "conv_tc_up2_allphase_kernel (Upsample x2 + Conv 128->64 3x3 fprop, folded, 4 TMEM accumulators)",
                    "bound": "tensor", "achieved": conv["tflops"], "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": conv["tflops"] / tf32_peak, "peak_source": peak_src + ", TF32 = bf16/2",
                    "ms_per_launch": conv["ms"], "algorithmic_gbs": conv["gbs"], "hbm_peak_gbs": hbm_peak,
                    "executed_gflop_per_launch": conv["flops"] / 1e9, "algorithmic_mbytes_per_launch": conv["bytes"] / 1e6,
                    "same_layer_dgrad_tflops": conv["dgrad_tflops"], "same_layer_wgrad_tflops": conv["wgrad_tflops"],
                    "cublas_tf32_gemm_8192_tflops": conv["cublas_tf32_gemm_tflops"],
                    "traffic": NCU_DRAM_TRAFFIC_BYTES_PER_LAUNCH}

    cpu = None
    if not args.no_cpu_baseline and world == 1:  # reported at N=1 only (rank 0); the scaling runs skip it
        ips, spstep, threads, batch = cpu_step_throughput(2, 1, budget_s=25.0)
        cpu = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"2 steps on {batch} of the {BATCH} images per step after 1 warm-up, oracle port of "
                         f"dcgan.py:146-183, stock torch CPU"}

    h2d = sum(t_.numel() * 4 for t_ in (host_imgs[0], host_z[0]))
    step_ms = ms / args.steps
    line = {
        "impl": args.impl,
        "metric": "DCGAN 64x64 images/sec (full G+D step)",
        "value": BATCH * world * args.steps / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "tf32 (fp32 storage, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": "DCGAN 64x64 synthetic, batch 128 per GPU (BASELINE configs[1])",
                   "global_batch": BATCH * world, "parallelism": f"dp{world}", "cuda_graph": not args.no_graph,
                   "l2": "per-step working set (~1.5 GB activations) exceeds the 126 MB L2; no explicit flush",
                   "algo": b200gan.Config.algo},
        "e2e": {"value": BATCH * world * args.steps / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": (calls_per_step or 0) * args.steps,
        "gpu_launches_note": "C-ABI launches of libb200gan kernels per step x steps (replayed from a CUDA graph)",
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "step_tensor_fraction_reference_form": GFLOP_PER_STEP_REFERENCE_FORM / step_ms / tf32_peak,
        "final_losses": {"g": losses[0], "d": losses[1]},
    }
    print(json.dumps(line), flush=True)
    _finish(world)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
