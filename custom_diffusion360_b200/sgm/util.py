"""Config factory of the drop-in boundary (mirrors the contract of the reference's sgm/util.py:
168-199 — `target:` dotted path + `params:` kwargs).  Selecting this implementation is done by
changing `target:` strings from `sgm.…` to `custom_diffusion360_b200.sgm.…` only."""
from __future__ import annotations

import importlib
from inspect import isfunction


def exists(x):
    return x is not None


def default(val, d):
    if val is not None:
        return val
    return d() if isfunction(d) else d


def get_obj_from_str(string: str, reload: bool = False):
    module, cls = string.rsplit(".", 1)
    mod = importlib.import_module(module)
    if reload:
        importlib.reload(mod)
    return getattr(mod, cls)


def instantiate_from_config(config):
    """`{"target": "pkg.mod.Class", "params": {...}}` -> Class(**params).  Accepts dicts or any
    mapping-like config object (OmegaConf DictConfig works unchanged)."""
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    params = config.get("params", dict()) if hasattr(config, "get") else dict()
    return get_obj_from_str(config["target"])(**params)


def append_dims(x, target_dims: int):
    """x[..., None, None] up to `target_dims` dims (sgm/util.py:192-199)."""
    extra = target_dims - x.ndim
    if extra < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * extra]


def append_zero(x):
    import torch

    return torch.cat([x, x.new_zeros([1])])
