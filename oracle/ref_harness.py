"""Import the UNMODIFIED reference modules in place from /root/reference — TEST INFRASTRUCTURE.

Used only here in the build container (the reference checkout does not exist on the GPU box) to
(1) validate oracle/sgm_oracle.py against the reference's own code and (2) generate the golden
vectors under tests/golden/.  Nothing is copied from the reference: the modules are imported from
where they lie, behind stand-ins for the third-party packages that are not installed
(omegaconf, xformers, pytorch3d, pytorch_lightning) — SURVEY.md §8c lists what each must provide.

The pytorch3d stand-in restates the documented PerspectiveCameras conventions (row vectors,
X_cam = X_world R + T, NDC) — pytorch3d itself is not available offline, so camera math is pinned
by hand-computed known answers only (tests/test_oracle_cameras.py).
"""
from __future__ import annotations

import collections
import os
import sys
import types

import torch
import torch.nn.functional as F

REF = os.environ.get("CD360_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "sgm", "modules"))


def _pkg(name, path=None):
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    sys.modules[name] = m
    return m


class StubCameras:
    """Minimal PerspectiveCameras: batched R [N,3,3], T [N,3], focal_length [N,2], principal_point [N,2]."""

    def __init__(self, R, T, focal_length, principal_point, image_size=None, device="cpu", **_):
        self.R = R.float()
        self.T = T.float()
        self.focal_length = focal_length.float()
        self.principal_point = principal_point.float()
        self.device = self.R.device

    def __len__(self):
        return self.R.shape[0]

    def __getitem__(self, i):
        if isinstance(i, int):
            if i >= len(self):
                raise IndexError  # the reference iterates `for cam in cam_batch`
            i = slice(i, i + 1)
        return StubCameras(self.R[i], self.T[i], self.focal_length[i], self.principal_point[i])

    def to(self, device):
        return StubCameras(self.R.to(device), self.T.to(device), self.focal_length.to(device),
                           self.principal_point.to(device))

    def get_camera_center(self):
        return -torch.einsum("nj,nkj->nk", self.T, self.R)

    def _view(self, p):
        if p.dim() == 2:
            p = p[None].expand(len(self), -1, -1)
        return torch.bmm(p, self.R) + self.T[:, None]

    def transform_points_ndc(self, p, **_):
        v = self._view(p)
        z = v[..., 2:3]
        return torch.cat([self.focal_length[:, None] * v[..., :2] / z + self.principal_point[:, None],
                          1 / z], -1)

    def unproject_points(self, xyd, world_coordinates=True, from_ndc=True):
        if xyd.dim() == 2:
            xyd = xyd[None].expand(len(self), -1, -1)
        z = xyd[..., 2:3]
        xy = (xyd[..., :2] - self.principal_point[:, None]) * z / self.focal_length[:, None]
        return torch.bmm(torch.cat([xy, z], -1) - self.T[:, None], self.R.transpose(1, 2))


def join_cameras_as_batch(cams):
    return StubCameras(*[torch.cat([getattr(c, f) for c in cams])
                         for f in ("R", "T", "focal_length", "principal_point")])


def cameras_from_packed(packed: torch.Tensor) -> StubCameras:
    """[n, 16] rows (R9 | T3 | f2 | pp2) -> stub camera batch (the reference's `pose` element)."""
    return StubCameras(packed[:, :9].reshape(-1, 3, 3), packed[:, 9:12], packed[:, 12:14], packed[:, 14:16])


_installed = False


def install() -> None:
    """Seed sys.modules so `sgm.modules.*` hot-path files import without their heavy dependencies."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF}")
    # package shells: skip sgm/__init__.py and sgm/modules/__init__.py (Lightning, kornia, open_clip)
    _pkg("sgm", os.path.join(REF, "sgm"))
    _pkg("sgm.modules", os.path.join(REF, "sgm", "modules"))
    _pkg("sgm.modules.diffusionmodules", os.path.join(REF, "sgm", "modules", "diffusionmodules"))

    class ListConfig(list):
        pass

    oc = _pkg("omegaconf")
    oc.ListConfig = ListConfig
    oc.OmegaConf = type("OmegaConf", (), {})
    _pkg("omegaconf.listconfig").ListConfig = ListConfig

    xo = _pkg("xformers.ops")
    _pkg("xformers").ops = xo
    # exact softmax attention on the [B*H, N, d] layout the reference passes (attention.py:394-408)
    xo.memory_efficient_attention = lambda q, k, v, attn_bias=None, op=None: \
        F.scaled_dot_product_attention(q, k, v)

    p3 = _pkg("pytorch3d")
    p3._C = types.SimpleNamespace(sample_pdf=None)  # unreachable in the shipped flow
    rend = _pkg("pytorch3d.renderer")
    rend.ray_bundle_to_ray_points = lambda rb: (
        rb.origins[..., None, :] + rb.lengths[..., :, None] * rb.directions[..., None, :])
    _pkg("pytorch3d.renderer.implicit")
    _pkg("pytorch3d.renderer.implicit.raysampling").RayBundle = collections.namedtuple(
        "RayBundle", "origins directions lengths xys")
    _pkg("pytorch3d.renderer.camera_utils").join_cameras_as_batch = join_cameras_as_batch
    _pkg("pytorch3d.renderer.cameras").PerspectiveCameras = StubCameras

    pl = _pkg("pytorch_lightning")
    pl.seed_everything = lambda s: torch.manual_seed(s)

    if not torch.cuda.is_available():
        # Raymarcher hard-codes device="cuda" (nerfsd_pytorch3d.py:249,251)
        _ls = torch.linspace
        torch.linspace = lambda *a, **k: _ls(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
    _installed = True


def import_reference():
    """Returns a namespace with the reference's own classes / functions for the hot path."""
    install()
    import importlib

    ns = types.SimpleNamespace()
    ns.openaimodel = importlib.import_module("sgm.modules.diffusionmodules.openaimodel")
    ns.attention = importlib.import_module("sgm.modules.attention")
    ns.nerf = importlib.import_module("sgm.modules.nerfsd_pytorch3d")
    ns.camray = importlib.import_module("sgm.modules.utils_cameraray")
    ns.util = importlib.import_module("sgm.modules.diffusionmodules.util")
    ns.discretizer = importlib.import_module("sgm.modules.diffusionmodules.discretizer")
    ns.denoiser_scaling = importlib.import_module("sgm.modules.diffusionmodules.denoiser_scaling")
    ns.guiders = importlib.import_module("sgm.modules.diffusionmodules.guiders")
    ns.sampling_utils = importlib.import_module("sgm.modules.diffusionmodules.sampling_utils")
    # sample.py: the two monkey-patched forwards used at inference (sample.py:33-136)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    cwd = os.getcwd()
    try:
        ns.sample = importlib.import_module("sample")
    finally:
        os.chdir(cwd)
    return ns


def build_reference_unet(ns, cfg: dict, state_dict: dict, patch_for_sampling: bool = True):
    """Instantiate the reference UNetModel for `cfg`, load `state_dict` (incl. `references`
    buffers), and apply sample.py's forward patches exactly as sample.py:247-270 does."""
    model = ns.openaimodel.UNetModel(**cfg)
    refs = {k: v for k, v in state_dict.items() if k.endswith("references")}
    params = {k: v for k, v in state_dict.items() if not k.endswith("references")}
    for name, module in model.named_modules():  # sgm/util.py:231-235
        if name.split(".")[-2:-1] == ["transformer_blocks"] and hasattr(module, "pose_emb_layers"):
            key = name + ".references"
            if key in refs:
                module.register_buffer("references", refs[key])
    missing, unexpected = model.load_state_dict(params, strict=False)
    missing = [m for m in missing if "raymarcher" not in m and "references" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    model.eval()
    if patch_for_sampling:
        for mod in model.modules():
            cname = mod.__class__.__name__
            if cname == "SpatialTransformer":
                mod.forward = ns.sample.customforward.__get__(mod, mod.__class__)
            elif cname == "BasicTransformerBlock":
                mod.forward = ns.sample._customforward.__get__(mod, mod.__class__)
    return model
