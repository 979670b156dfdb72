"""Camera packing and ray grids for FeatureNeRF (host side).

The reference walks Python lists of pytorch3d `PerspectiveCameras` and loops per view
(sgm/modules/utils_cameraray.py:61-196, 245-314).  Here the cameras are packed once into one
fp32 tensor [b, n+1, 16] = R (9, row-major; PyTorch3D row-vector convention X_cam = X_world R + T)
| T (3) | focal (2) | principal point (2), index 0 = target view, and all per-ray geometry runs
inside the CUDA kernels (csrc/nerf.cu).
"""
from __future__ import annotations

from typing import Sequence

import torch


def pack_camera_batch(cam) -> torch.Tensor:
    """One pytorch3d-like camera batch (fields R [m,3,3], T [m,3], focal_length [m,2] or [m,1],
    principal_point [m,2]) or an already packed [m,16] tensor -> fp32 [m, 16] on the host."""
    if isinstance(cam, torch.Tensor):
        assert cam.shape[-1] == 16
        return cam.detach().float().cpu()
    R = torch.as_tensor(cam.R).detach().float().cpu().reshape(-1, 9)
    m = R.shape[0]
    T = torch.as_tensor(cam.T).detach().float().cpu().reshape(m, 3)
    f = torch.as_tensor(cam.focal_length).detach().float().cpu().reshape(m, -1)
    if f.shape[1] == 1:
        f = f.expand(m, 2)
    pp = torch.as_tensor(cam.principal_point).detach().float().cpu().reshape(m, 2)
    return torch.cat([R, T, f, pp], dim=1)


def pack_pose(pose: Sequence, device) -> torch.Tensor:
    """`pose` as the reference passes it (python list, one camera batch of n+1 cameras per UNet
    batch row; sample.py:302,326; data_co3d.py:631) -> fp32 [b, n+1, 16] on `device`."""
    if isinstance(pose, torch.Tensor):
        return pose.to(device=device, dtype=torch.float32).contiguous()
    cache = {}
    rows = []
    for cam in pose:  # CFG repeats the same object (`pose * 3`, sample.py:169): pack it once
        key = id(cam)
        if key not in cache:
            cache[key] = pack_camera_batch(cam)
        rows.append(cache[key])
    return torch.stack(rows).to(device).contiguous()


_XY_CACHE: dict = {}


def patch_ray_xy(res: int, device) -> torch.Tensor:
    """NDC centres of a res x res patch grid, row-major, [res*res, 2] — the deterministic branch of
    get_patch_raybundle (utils_cameraray.py:106-153): midpoints of linspace(1, -1, res+1),
    meshgrid(indexing='xy')."""
    key = (res, str(device))
    hit = _XY_CACHE.get(key)
    if hit is not None:       # (cached: a host -> device copy is not allowed while a CUDA graph is captured)
        return hit
    edges = torch.linspace(1, -1, res + 1, dtype=torch.float32)
    centers = (edges[:-1] + edges[1:]) / 2
    xs = centers[None, :].expand(res, res)  # x varies along the fast (column) index
    ys = centers[:, None].expand(res, res)
    out = torch.stack([xs.reshape(-1), ys.reshape(-1)], dim=-1).contiguous().to(device)
    _XY_CACHE[key] = out
    return out
