"""Training steps of the hot-path scripts, re-enacted around the drop-in modules, plus the CUDA-graph
runner used by bench.py.  The step bodies follow the reference loops line by line so that the
unmodified scripts (via launch.py) and this module execute the same sequence of operator calls.
"""
import contextlib

import torch


@contextlib.contextmanager
def frozen(module):
    """Temporarily mark a network's parameters as not requiring grad.  The reference back-propagates the
    generator loss into the discriminator's weights too (dcgan.py:168) and throws those gradients away at
    optimizer_D.zero_grad() (dcgan.py:175); skipping that dead weight-gradient work changes no result."""
    params = [p for p in module.parameters() if p.requires_grad]
    for p in params:
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p in params:
            p.requires_grad_(True)


def dcgan_step(generator, discriminator, opt_g, opt_d, real_imgs, z, loss=None, valid=None, fake=None,
               reduce_g=None, reduce_d=None, skip_dead_wgrad=True):
    """implementations/dcgan/dcgan.py:146-183.  `reduce_*`: optional gradient all-reduce hooks invoked
    right before the corresponding optimizer step (data parallel)."""
    loss = loss or torch.nn.BCELoss()
    n = real_imgs.shape[0]
    if valid is None:
        valid = torch.ones(n, 1, device=real_imgs.device)   # dcgan.py:147
        fake = torch.zeros(n, 1, device=real_imgs.device)   # dcgan.py:148
    opt_g.zero_grad()                                        # :157
    gen_imgs = generator(z)                                  # :163
    with frozen(discriminator) if skip_dead_wgrad else contextlib.nullcontext():
        g_loss = loss(discriminator(gen_imgs), valid)        # :166
        g_loss.backward()                                    # :168
    if reduce_g is not None:
        reduce_g()
    opt_g.step()                                             # :169
    opt_d.zero_grad()                                        # :175
    real_loss = loss(discriminator(real_imgs), valid)        # :178
    fake_loss = loss(discriminator(gen_imgs.detach()), fake) # :179
    d_loss = (real_loss + fake_loss) / 2                     # :180
    d_loss.backward()                                        # :182
    if reduce_d is not None:
        reduce_d()
    opt_d.step()                                             # :183
    return g_loss.detach(), d_loss.detach(), gen_imgs.detach()


class GraphedStep:
    """Capture one training step into a CUDA graph (static input/output buffers) and replay it.
    The step function must be free of host synchronisation; optimizers must be `capturable`."""

    def __init__(self, step_fn, example_inputs, warmup=3):
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.static_out = step_fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = step_fn(*self.static_in)

    def __call__(self, *inputs):
        for s, t in zip(self.static_in, inputs):
            s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_out
