"""CPU emulation of what a TF32 convolution does to gradients behind a LeakyReLU (no GPU needed).

Upsample -> Conv3x3 -> BatchNorm(eps 0.8) -> LeakyReLU(slope) in fp32, against the same block whose convolution
operands are cut to 10 mantissa bits the way the tensor cores do it (activations truncated, weights rounded to
nearest).  With slope 1.0 the gradients lose ~4e-4 like the forward; with slope 0.2 they lose 3e-3 .. 1.3e-2,
because a few pre-activations land on the other side of the kink.  tests/test_gpu_ops.py cites these numbers for its
gradient bounds.
"""
import torch


def trunc(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def rn(t):
    return ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


class Tf32Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return torch.nn.functional.conv2d(trunc(x), rn(w), b, 1, 1)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = torch.nn.grad.conv2d_input(x.shape, rn(w), trunc(dy), 1, 1)
        dw = torch.nn.grad.conv2d_weight(trunc(x), w.shape, trunc(dy), 1, 1)
        return dx, dw, dy.sum((0, 2, 3))


if __name__ == "__main__":
    torch.manual_seed(11)
    for slope in (0.2, 1.0):
        for cin, cout, h, w, n in [(64, 64, 6, 5, 3), (128, 64, 16, 16, 2)]:
            blk = torch.nn.Sequential(torch.nn.Upsample(scale_factor=2), torch.nn.Conv2d(cin, cout, 3, 1, 1),
                                      torch.nn.BatchNorm2d(cout, 0.8), torch.nn.LeakyReLU(slope))
            x, gy = torch.randn(n, cin, h, w), torch.randn(n, cout, 2 * h, 2 * w)
            xi = x.clone().requires_grad_(True)
            y = blk(xi)
            y.backward(gy)
            gx, gw = xi.grad.clone(), blk[1].weight.grad.clone()
            blk.zero_grad()
            xi = x.clone().requires_grad_(True)
            y2 = blk[3](blk[2](Tf32Conv.apply(blk[0](xi), blk[1].weight, blk[1].bias)))
            y2.backward(gy)
            print(f"slope {slope} {cin}->{cout} {h}x{w} n{n}: fwd {rel(y2, y):.2e}  dx {rel(xi.grad, gx):.2e}  "
                  f"dw {rel(blk[1].weight.grad, gw):.2e}")
