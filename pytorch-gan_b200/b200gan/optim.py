"""torch.optim.Adam for the hot path (dcgan.py:134-135, pix2pix.py:80-81, cyclegan.py:87-91, wgan_gp.py:106-107):
same constructor, same arithmetic (torch's _single_tensor_adam, including where Python doubles meet fp32), but ONE
kernel launch per step() for all parameters (b200gan_adam_multi) instead of ~10 multi-tensor passes, and a step
count that lives on the device, so a training step can be captured in a CUDA graph without `capturable=True`.

`grad_scale` folds the 1/world_size of a data-parallel all-reduce(sum) into the update (b200gan/ddp.py)."""
import ctypes

import torch

from . import _lib


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **unsupported):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("b200gan.optim.Adam: weight_decay / amsgrad are not on the hot path")
        for k, v in unsupported.items():
            if k in ("foreach", "capturable", "fused", "differentiable", "maximize") and not v:
                continue
            if k == "capturable":  # the device-side step count makes every step capturable
                continue
            raise NotImplementedError(f"b200gan.optim.Adam: option {k}={v!r}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.grad_scale = 1.0
        self.grad_override = None  # {param: flat gradient tensor} set by ddp.GradReducer for one step

    def _group_state(self, group):
        st = group.get("_b200")
        if st is None:
            ps = [p for p in group["params"] if p.requires_grad]
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("b200gan.optim.Adam: parameters must be contiguous fp32 CUDA tensors "
                                       "(no CPU fallback)")
            dev = ps[0].device if ps else torch.device("cuda")
            total = sum(p.numel() for p in ps)
            flat_m = torch.zeros(total, device=dev, dtype=torch.float32)
            flat_v = torch.zeros(total, device=dev, dtype=torch.float32)
            step = torch.zeros(2, device=dev, dtype=torch.float32)
            off = 0
            for p in ps:
                n = p.numel()
                # the usual torch state layout, so state_dict() / load_state_dict() keep working
                self.state[p] = {"step": step[0:1].view(()), "exp_avg": flat_m[off:off + n].view_as(p),
                                 "exp_avg_sq": flat_v[off:off + n].view_as(p)}
                off += n
            st = {"step": step, "params": ps}
            group["_b200"] = st
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            st = self._group_state(group)
            rows = []
            for p in st["params"]:
                g = p.grad if self.grad_override is None else self.grad_override.get(p, p.grad)
                if g is None:
                    continue
                if not g.is_contiguous():
                    g = g.contiguous()
                s = self.state[p]
                rows.append((p.data_ptr(), g.data_ptr(), s["exp_avg"].data_ptr(), s["exp_avg_sq"].data_ptr(), p.numel(), g))
            if not rows:
                continue
            table = (_lib.AdamTensor * len(rows))()
            for i, r in enumerate(rows):
                table[i].p, table[i].g, table[i].m, table[i].v, table[i].n = r[0], r[1], r[2], r[3], r[4]
            b1, b2 = group["betas"]
            _lib.check(lib.b200gan_adam_multi(table, len(rows), float(group["lr"]), float(b1), float(b2),
                                              float(group["eps"]), float(self.grad_scale), st["step"].data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), "adam_multi")
            # the kernel wrote the parameters behind autograd's back: bump their version counters, which is what
            # invalidates the packed-weight caches of the conv modules (functional.PackCache) -- no kernel involved
            updated = [p for p in st["params"] if p.grad is not None or (self.grad_override and p in self.grad_override)]
            torch.autograd.graph.increment_version(updated)
            # ... and refresh every packed copy of the updated conv weights in one launch
            from . import functional
            functional.refresh_packs([p for p in updated if p.dim() == 4])
        return loss

    def state_dict(self):
        sd = super().state_dict()
        for g in sd["param_groups"]:
            g.pop("_b200", None)
        return sd
