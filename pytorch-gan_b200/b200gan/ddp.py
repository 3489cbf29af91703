"""Data-parallel gradient exchange: one process per GPU, batch sharded across
ranks, one NCCL all-reduce (sum, then 1/world) of a network's gradients right before its optimizer
step -- the only collective on the path.  BatchNorm statistics stay per replica (torch-DDP semantics;
the reference has no SyncBN)."""
import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


class GradReducer:
    def __init__(self, params, world):
        self.params = [p for p in params if p.requires_grad]
        self.world = world

    def __call__(self):
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads or self.world == 1:
            return
        flat = _flatten_dense_tensors(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(self.world)
        for g, f in zip(grads, _unflatten_dense_tensors(flat, grads)):
            g.copy_(f)
