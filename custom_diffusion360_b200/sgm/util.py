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
    cond = getattr(engine, "conditioner", None)
    if cond is not None and hasattr(cond, "embedders"):
        # conditioner.embedders.{0,1}.*: the base checkpoint's token embeddings have one row less than a
        # tower built with `modifier_token` — the reference concatenates the delta checkpoint's `embed`
        # rows behind them (sgm/util.py:216-228)
        csd = {k[len("conditioner."):]: v for k, v in base_sd.items() if k.startswith("conditioner.")}
        own = cond.state_dict()
        for k in list(csd):
            if k in own and k.endswith("token_embedding.weight") and csd[k].shape[0] < own[k].shape[0]:
                full = own[k].clone()
                full[: csd[k].shape[0]] = csd[k]
                csd[k] = full
        cm, cu = cond.load_state_dict(csd, strict=False)
        for m_ in cond.modules():
            if hasattr(m_, "invalidate_packed"):
                m_.invalidate_packed()
        if delta_sd is not None and "embed" in delta_sd and hasattr(cond, "load_modifier_token_rows"):
            cond.load_modifier_token_rows(delta_sd["embed"])
        missing = missing + ["conditioner." + k for k in cm if not k.endswith("logit_scale")]
        unexpected = unexpected + ["conditioner." + k for k in cu if "position_ids" not in k and "attn_mask" not in k]
    if verbose:
        print(f"missing: {missing}\nunexpected: {unexpected}")
    return missing, unexpected


def append_zero(x):
    import torch

    return torch.cat([x, x.new_zeros([1])])


# ---- delta checkpoint + camera.bin I/O without pytorch3d -------------------------
def delta_state_dict(engine, embed=None) -> dict:
    """What the reference's `CUDACallback.on_save_checkpoint` keeps (main.py:611-625): every
    state-dict entry whose key contains `pose` (but not `raymarcher`) or `references`, under the
    Lightning module's key names (`model.diffusion_model.<name>`), plus `embed` — the trained
    `<new1>` token rows of the two text encoders, `[clip_l_row, open_clip_row]`.  `embed` defaults to
    the engine's own conditioner (`engine.conditioner.modifier_token_rows()`) when it has one."""
    st = engine.state_dict()
    out = {k: v.detach().clone() for k, v in st.items()
           if ("pose" in k and "raymarcher" not in k) or "references" in k}
    if embed is None:
        cond = getattr(engine, "conditioner", None)
        if cond is not None and hasattr(cond, "modifier_token_rows"):
            embed = cond.modifier_token_rows()
    if embed is not None:
        out["embed"] = [e.detach().clone() for e in embed]
    return out


def save_delta_checkpoint(engine, path, embed=None, **extra):
    """`torch.save({'delta_state_dict': …})` — the file `sample.py --custom_model_dir` reads back
    through `load_model_from_config`'s delta branch (sgm/util.py:225-237).  That loader indexes
    `sd_delta['embed'][0]` and `[1]` unconditionally, so a checkpoint without the two token rows is
    unreadable by the reference: `embed` is required (from the argument or the engine's conditioner)."""
    import torch

    delta = delta_state_dict(engine, embed)
    if "embed" not in delta or len(delta["embed"]) != 2:
        raise ValueError("save_delta_checkpoint needs `embed` = [clip_l_token_row, open_clip_token_row] "
                         "(the reference loader reads sd_delta['embed'][0] and [1], sgm/util.py:225-228); "
                         "pass it or attach a conditioner that provides modifier_token_rows()")
    torch.save({"delta_state_dict": delta, **extra}, path)


class _CameraShell:
    """Stand-in for any pytorch3d class met while unpickling `camera.bin`: keeps the pickled
    attribute dict (R, T, focal_length, principal_point, …) and nothing else."""

    def __init__(self, *a, **kw):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {})


def _camera_unpickler():
    import pickle
    import types

    class Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.split(".")[0] == "pytorch3d":
                return type(name, (_CameraShell,), {})
            return super().find_class(module, name)

    return types.SimpleNamespace(Unpickler=Unpickler, load=lambda f, **kw: Unpickler(f, **kw).load(),
                                 __name__="pickle")


def _as_packed(cam):
    import torch

    from .modules.utils_cameraray import pack_camera_batch

    if isinstance(cam, torch.Tensor):
        return pack_camera_batch(cam)
    d = cam if isinstance(cam, dict) else cam.__dict__
    # nn.Module pickles keep plain tensor attributes in __dict__ and registered ones in _buffers / _parameters
    fields = {}
    for key in ("R", "T", "focal_length", "principal_point"):
        v = d.get(key)
        if v is None:
            for bag in ("_buffers", "_parameters"):
                v = (d.get(bag) or {}).get(key, v)
        if v is None:
            raise KeyError(f"camera.bin entry has no `{key}`")
        fields[key] = v
    import types

    return pack_camera_batch(types.SimpleNamespace(**fields))


def load_camera_bin(path):
    """`camera.bin` = `torch.save([cameras_val, cameras_train])`, two lists of pytorch3d
    `PerspectiveCameras` (main.py:1025-1029, read at sample.py:273).  Returns the same pair as packed
    fp32 tensors `[N, 16]` = R(9) | T(3) | focal(2) | principal point(2), one row per camera —
    directly usable as `pose` entries (`torch.cat([target_row, train_rows[choices]])`) — without
    importing pytorch3d: its classes are unpickled into attribute shells."""
    import torch

    with open(path, "rb") as f:
        cams_val, cams_train = torch.load(f, map_location="cpu", pickle_module=_camera_unpickler(),
                                          weights_only=False)

    def rows(lst):
        if isinstance(lst, torch.Tensor):
            return lst.float()
        if not isinstance(lst, (list, tuple)):
            lst = [lst]
        return torch.cat([_as_packed(c) for c in lst], dim=0)

    return rows(cams_val), rows(cams_train)


def reference_choices(n_train: int, num_ref: int = 8):
    """The stored reference views sample.py conditions on (sample.py:275-278)."""
    import torch

    max_diff = n_train / num_ref
    return [int(x) for x in torch.linspace(0, n_train - max_diff, num_ref)]


def sample_pose(cams_val, cams_train, target_index: int, choices=None):
    """One `pose` entry as sample.py builds it (:302,326): target camera followed by the chosen
    training cameras, packed `[1 + len(choices), 16]`."""
    import torch

    choices = reference_choices(cams_train.shape[0]) if choices is None else choices
    return torch.cat([cams_val[target_index:target_index + 1], cams_train[list(choices)]], dim=0)
