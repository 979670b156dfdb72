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


def load_checkpoints(engine, base_sd: dict, delta_sd: dict = None, verbose: bool = False):
    """Checkpoint key contract of the reference (sgm/util.py:202-251, main.py:611-625):
    `base_sd` uses `model.diffusion_model.<name>` keys (sd_xl_base_1.0.safetensors; conditioner /
    first-stage keys are ignored here), `delta_sd` = checkpoint['delta_state_dict'] with the pose
    weights, the per-block `references` buffers and `embed` (token rows, not ours to load).
    Returns (missing, unexpected) for the UNet part."""
    prefix = "model.diffusion_model."
    unet = engine.model.diffusion_model
    sd = {k[len(prefix):]: v for k, v in base_sd.items() if k.startswith(prefix)}
    refs = {}
    if delta_sd is not None:
        for k, v in delta_sd.items():
            if not k.startswith(prefix):
                continue
            name = k[len(prefix):]
            if name.endswith(".references"):
                refs[name] = v
            else:
                sd[name] = v
    missing, unexpected = unet.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "raymarcher" not in m]
    if refs:
        dev = next(unet.parameters()).device
        unet.register_references({k: v.to(dev) for k, v in refs.items()})
    if verbose:
        print(f"missing: {missing}\nunexpected: {unexpected}")
    return missing, unexpected


def append_zero(x):
    import torch

    return torch.cat([x, x.new_zeros([1])])
