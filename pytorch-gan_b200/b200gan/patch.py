"""Rebind the torch.nn leaf classes the reference scripts look up by attribute (`nn.Conv2d(...)`,
dcgan.py:13,55) to the b200gan drop-ins, so an implementations/*/*.py script runs unmodified."""
import contextlib

import torch.nn as tnn

_saved = {}


def patch():
    from .nn import REPLACEMENTS
    if _saved:
        return
    for name, cls in REPLACEMENTS.items():
        _saved[name] = getattr(tnn, name)
        setattr(tnn, name, cls)


def unpatch():
    for name, cls in _saved.items():
        setattr(tnn, name, cls)
    _saved.clear()


@contextlib.contextmanager
def patched():
    patch()
    try:
        yield
    finally:
        unpatch()
