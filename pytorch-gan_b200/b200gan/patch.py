"""Rebind the torch.nn leaf classes the reference scripts look up by attribute (`nn.Conv2d(...)`,
dcgan.py:13,55) to the b200gan drop-ins, so an implementations/*/*.py script runs unmodified; and
`torch.optim.Adam` (dcgan.py:134-135) to the one-launch multi-tensor Adam when every parameter is a CUDA fp32 tensor
(anything else -- e.g. gan.py on CPU, BASELINE config 0 -- gets the stock optimizer)."""
import contextlib

import torch
import torch.nn as tnn

_saved = {}
_saved_adam = []


def _adam_dispatch(stock_adam):
    class Adam(stock_adam):
        def __new__(cls, params, *args, **kw):
            from . import optim
            params = list(params)
            flat = [p for g in params for p in (g["params"] if isinstance(g, dict) else [g])]
            plain = not kw.get("weight_decay") and not kw.get("amsgrad") and len(args) <= 3
            if flat and plain and all(torch.is_tensor(p) and p.is_cuda and p.dtype == torch.float32 for p in flat):
                try:
                    return optim.Adam(params, *args, **kw)
                except NotImplementedError:
                    pass
            return stock_adam(params, *args, **kw)
    Adam.__name__ = Adam.__qualname__ = "Adam"
    return Adam


def patch(optimizers=True):
    from .nn import REPLACEMENTS
    if _saved:
        return
    for name, cls in REPLACEMENTS.items():
        _saved[name] = getattr(tnn, name)
        setattr(tnn, name, cls)
    if optimizers:
        _saved_adam.append(torch.optim.Adam)
        torch.optim.Adam = _adam_dispatch(torch.optim.Adam)
    # name-based init through Module.apply writes parameters through .data (dcgan.py:36-42): packed-weight caches key on
    # the version counter, which .data writes do not bump -- drop them whenever a module tree is re-initialised
    _saved["__apply__"] = tnn.Module.apply

    def apply(self, fn):
        out = _saved["__apply__"](self, fn)
        for m in self.modules():
            m.__dict__.pop("_b200_cache", None)
        return out
    tnn.Module.apply = apply


def unpatch():
    for name, cls in _saved.items():
        if name == "__apply__":
            tnn.Module.apply = cls
        else:
            setattr(tnn, name, cls)
    _saved.clear()
    if _saved_adam:
        torch.optim.Adam = _saved_adam.pop()


@contextlib.contextmanager
def patched():
    patch()
    try:
        yield
    finally:
        unpatch()
