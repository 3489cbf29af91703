"""Stand-alone CUDA-event timings of the CycleGAN step's building blocks at the BASELINE size (batch 8, 256x256,
cyclegan/models.py): the 7x7 reflection-padded stem and output layer, a residual-block conv with its reflection pad and
InstanceNorm, through the drop-in modules (forward + backward of one module each).  A one-minute substitute for the
12-minute ncu launch list of the whole step.
    python tools/cyclegan_ops.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pytorch-gan_b200"))
from b200gan import nn as bnn  # noqa: E402


def fwd_bwd_ms(mod, x, iters=5, need_dx=True):
    x = x.clone().requires_grad_(need_dx)
    y = mod(x)
    gy = torch.randn_like(y)
    for _ in range(2):
        mod.zero_grad(set_to_none=True)
        y = mod(x)
        y.backward(gy)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(iters):
        mod.zero_grad(set_to_none=True)
        e[0].record()
        y = mod(x)
        e[1].record()
        y.backward(gy)
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1])
        tb += e[1].elapsed_time(e[2])
    return tf / iters * 1e3, tb / iters * 1e3


def main():
    n = 8
    cases = [
        ("stem: ReflectionPad2d(3) Conv2d(3,64,7) IN ReLU  (input = data)", bnn.Sequential(
            bnn.ReflectionPad2d(3), bnn.Conv2d(3, 64, 7), bnn.InstanceNorm2d(64), bnn.ReLU(inplace=True)), (n, 3, 256, 256), False),
        ("stem, input = generated image (needs dx)", bnn.Sequential(
            bnn.ReflectionPad2d(3), bnn.Conv2d(3, 64, 7), bnn.InstanceNorm2d(64), bnn.ReLU(inplace=True)), (n, 3, 256, 256), True),
        ("output: ReflectionPad2d(3) Conv2d(64,3,7) Tanh", bnn.Sequential(
            bnn.ReflectionPad2d(3), bnn.Conv2d(64, 3, 7), bnn.Tanh()), (n, 64, 256, 256), True),
        ("residual half: ReflectionPad2d(1) Conv2d(256,256,3) IN ReLU", bnn.Sequential(
            bnn.ReflectionPad2d(1), bnn.Conv2d(256, 256, 3), bnn.InstanceNorm2d(256), bnn.ReLU(inplace=True)), (n, 256, 64, 64), True),
        ("down: Conv2d(64,128,3,2,1) IN ReLU", bnn.Sequential(
            bnn.Conv2d(64, 128, 3, stride=2, padding=1), bnn.InstanceNorm2d(128), bnn.ReLU(inplace=True)), (n, 64, 256, 256), True),
        ("up: Upsample(2) Conv2d(256,128,3,1,1) IN ReLU", bnn.Sequential(
            bnn.Upsample(scale_factor=2), bnn.Conv2d(256, 128, 3, stride=1, padding=1), bnn.InstanceNorm2d(128),
            bnn.ReLU(inplace=True)), (n, 256, 64, 64), True),
        ("D block1: Conv2d(3,64,4,2,1) LeakyReLU", bnn.Sequential(
            bnn.Conv2d(3, 64, 4, stride=2, padding=1), bnn.LeakyReLU(0.2, inplace=True)), (n, 3, 256, 256), True),
        ("D block2: Conv2d(64,128,4,2,1) IN LeakyReLU", bnn.Sequential(
            bnn.Conv2d(64, 128, 4, stride=2, padding=1), bnn.InstanceNorm2d(128), bnn.LeakyReLU(0.2, inplace=True)),
         (n, 64, 128, 128), True),
    ]
    print(f"{'module (batch 8)':<70}{'fwd us':>9}{'bwd us':>9}")
    for name, mod, shape, need_dx in cases:
        mod = mod.cuda().train()
        x = torch.randn(*shape, device="cuda").contiguous(memory_format=torch.channels_last)
        tf, tb = fwd_bwd_ms(mod, x, need_dx=need_dx)
        print(f"{name:<70}{tf:9.0f}{tb:9.0f}", flush=True)


if __name__ == "__main__":
    main()
