"""Data-parallel gradient exchange: one process per GPU, batch sharded across ranks, one NCCL
all-reduce(sum) of a network's gradients right before its optimizer step -- the only collective on the path
(dcgan.py:169,183; pix2pix.py:152,172; cyclegan.py:205,222,239).  BatchNorm statistics stay per replica (torch-DDP
semantics; the reference has no SyncBN).

The reducer gathers the gradients into ONE flat bucket (one kernel), reduces the bucket (one NCCL call) and hands the
optimizer views of the bucket: no unflatten copies, and the 1/world average is folded into the Adam kernel's
grad_scale.  With `overlap`, bucket + all-reduce + optimizer step run on a side stream while the main stream goes on
with work that does not depend on the updated weights (the whole discriminator phase of a GAN step, for the generator's
bucket); `join()` at the end of the step orders the streams again.  The side stream forks and joins inside a CUDA-graph
capture like any other stream."""
import os

import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, params, world, opt=None, overlap=None):
        self.params = [p for p in params if p.requires_grad]
        self.world = world
        self.opt = opt  # a b200gan.optim.Adam (reads the bucket directly) or None (gradients are written back)
        if overlap is None:
            overlap = os.environ.get("B200GAN_DDP_OVERLAP", "1") not in ("", "0")
        self.overlap = bool(overlap) and world > 1
        self.side = None
        self._keep = None

    def _reduce(self):
        ps = [p for p in self.params if p.grad is not None]
        if not ps:
            return ps, None
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
        if self.world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return ps, flat

    def __call__(self):
        """Blocking form for a foreign optimizer: gradients are averaged in place."""
        if self.world == 1:
            return
        ps, flat = self._reduce()
        if flat is None:
            return
        flat.div_(self.world)
        off = 0
        for p in ps:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n

    def reduce_and_step(self, opt):
        """All-reduce this network's gradients and run its optimizer step (possibly on the side stream)."""
        if self.world == 1:
            opt.step()
            return
        if opt is not self.opt or self.opt is None:
            self()
            opt.step()
            return
        main = torch.cuda.current_stream()
        if self.overlap:
            if self.side is None:
                self.side = torch.cuda.Stream()
            self.side.wait_stream(main)
            ctx = torch.cuda.stream(self.side)
        else:
            ctx = torch.cuda.stream(main)
        with ctx:
            ps, flat = self._reduce()
            if flat is not None:
                over, off = {}, 0
                for p in ps:
                    n = p.numel()
                    over[p] = flat[off:off + n]
                    off += n
                opt.grad_override, opt.grad_scale = over, 1.0 / self.world
                opt.step()
                opt.grad_override, opt.grad_scale = None, 1.0
                self._keep = (flat, [p.grad for p in ps])  # alive until the streams are joined

    def join(self):
        if self.overlap and self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        self._keep = None
