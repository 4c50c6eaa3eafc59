"""Parameter holders: nn.Module trees that carry the reference's state_dict keys so that
`load_state_dict`, `.half()`, `.to(device)`, `parameters()` behave as on the reference model, while the
arithmetic lives in libgigaam_b200.so."""
from __future__ import annotations

import weakref
from typing import Dict, Iterable, Optional, Tuple

import torch
from torch import nn


class Holder(nn.Module):
    """Container without behaviour; only gives dotted state_dict keys a module tree to live in."""


_BUFFER_LEAVES = ("running_mean", "running_var", "num_batches_tracked", "window", "fb")


def attach(root: nn.Module, key: str, tensor: torch.Tensor) -> None:
    parts = key.split(".")
    m = root
    for p in parts[:-1]:
        if p not in m._modules:
            m.add_module(p, Holder())
        m = m._modules[p]
    if parts[-1] in _BUFFER_LEAVES:
        m.register_buffer(parts[-1], tensor)
    else:
        m.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


def build_tree(root: nn.Module, entries: Iterable[Tuple[str, torch.Tensor]], strip_prefix: str) -> None:
    for key, t in entries:
        assert key.startswith(strip_prefix), key
        attach(root, key[len(strip_prefix):], t)


class Bound(nn.Module):
    """A component whose compute is served by the engine of the model that owns it."""

    def __init__(self):
        super().__init__()
        self.__dict__["_owner_ref"] = None

    def _bind(self, owner) -> None:
        self.__dict__["_owner_ref"] = weakref.ref(owner)

    def _engine(self):
        ref = self.__dict__.get("_owner_ref")
        owner = ref() if ref is not None else None
        if owner is None:
            raise RuntimeError(
                f"{type(self).__name__} is not bound to a model: build it through gigaam_b200.GigaAM / GigaAMASR / "
                "load_model(); the kernels need the whole model's weights on one device (no CPU path exists)")
        return owner._get_engine()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        ref = self.__dict__.get("_owner_ref")
        owner = ref() if ref is not None else None
        if owner is not None:
            owner._invalidate_engine()
        return out
