"""Reading reference checkpoints without omegaconf.

A reference `.ckpt` is a `torch.save` of `{"cfg": omegaconf.DictConfig, "state_dict": ...}` (gigaam/__init__.py:167).
Unpickling the cfg needs the omegaconf classes (`omegaconf.dictconfig.DictConfig`, `omegaconf.nodes.*`,
`omegaconf.base.ContainerMetadata` ...), which are absent on the build / bench boxes.  The unpickler below resolves
every global under `omegaconf.` to a permissive stand-in that only records its pickled state; afterwards the tree is
converted to plain dicts / lists / scalars, which is all `model.py` reads from a cfg.  When omegaconf is installed the
ordinary `torch.load` is used and the cfg is left as it is."""
from __future__ import annotations

import pickle
import warnings
from typing import Any, Dict

import torch


class _Stub:
    """Stand-in for any omegaconf class: keeps constructor arguments and pickled state, interprets nothing."""
    _qualname = "omegaconf.?"

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        self.__dict__["_args"] = args
        self.__dict__.update(kwargs)

    def __setstate__(self, state: Any) -> None:
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):   # (dict state, slots state)
            merged = dict(state[0] or {})
            merged.update(state[1])
            state = merged
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state

    def __call__(self, *args: Any, **kwargs: Any) -> "_Stub":   # enum-like lookups pickled as Class(value)
        return self


_STUB_CACHE: Dict[str, type] = {}


def _stub_class(module: str, name: str) -> type:
    key = f"{module}.{name}"
    if key not in _STUB_CACHE:
        _STUB_CACHE[key] = type(name, (_Stub,), {"_qualname": key})
    return _STUB_CACHE[key]


class _Unpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        if module == "omegaconf" or module.startswith("omegaconf."):
            return _stub_class(module, name)
        return super().find_class(module, name)


class _PickleModule:
    """What `torch.load(pickle_module=...)` expects: a module-like object with Unpickler / load / loads."""
    __name__ = "gigaam_b200.ckpt"
    Unpickler = _Unpickler
    UnpicklingError = pickle.UnpicklingError
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL

    @staticmethod
    def load(f, **kwargs):
        return _Unpickler(f, **kwargs).load()

    @staticmethod
    def loads(data, **kwargs):
        import io
        return _Unpickler(io.BytesIO(data), **kwargs).load()


def to_plain(obj: Any) -> Any:
    """omegaconf stand-ins -> plain containers: containers keep their `_content`, value nodes their `_val`."""
    if isinstance(obj, _Stub):
        d = obj.__dict__
        if "_content" in d:
            return to_plain(d["_content"])
        if "_val" in d:
            return to_plain(d["_val"])
        return None          # metadata / flags objects carry nothing the model reads
    if isinstance(obj, dict):
        return {to_plain(k) if isinstance(k, _Stub) else k: to_plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [to_plain(v) for v in obj]
    return obj


def load_checkpoint(path: str) -> Dict:
    """`torch.load` of a reference checkpoint; the cfg comes back as plain containers when omegaconf is missing."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=FutureWarning)
        try:
            import omegaconf  # noqa: F401
            return torch.load(path, map_location="cpu", weights_only=False)
        except ImportError:
            pass
        ck = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_PickleModule)
    for key in ("cfg", "hyper_parameters"):
        if key in ck:
            ck[key] = to_plain(ck[key])
    return ck
