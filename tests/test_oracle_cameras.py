"""Known answers for the camera conventions restated from PyTorch3D's documentation (pytorch3d is
a pinned but un-vendored dependency of the reference and is not installable offline, so this is
all the pinning there is for the camera math: SURVEY.md §8c "parity unpinned")."""
import math

import torch

from oracle import sgm_oracle as O


def _cam(R, T, f=2.0):
    return O.pack_cameras(torch.tensor(R, dtype=torch.float32)[None], torch.tensor(T, dtype=torch.float32)[None],
                          torch.tensor([[f, f]]), torch.zeros(1, 2))[0]


def test_axis_point_projects_to_ndc_origin():
    # camera at (0,0,-3) looking down +Z: R = I, T = -C R = (0,0,3)
    cam = _cam(torch.eye(3).tolist(), [0.0, 0.0, 3.0])
    assert torch.allclose(O.camera_centers(cam), torch.tensor([0.0, 0.0, -3.0]))
    xy = O.transform_points_ndc(cam, torch.tensor([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 1.0]]))
    assert torch.allclose(xy[0], torch.zeros(2))
    # +X world -> +x NDC ("left" in PyTorch3D's NDC), x = f * X / Z = 2 * 1 / 3
    assert torch.allclose(xy[1], torch.tensor([2.0 / 3.0, 0.0]))
    assert torch.allclose(xy[2], torch.tensor([0.0, 2.0 * 1.0 / 4.0]))


def test_unproject_is_inverse_of_project():
    cams = O.lookat_cameras(4, seed=5)
    xy = O.patch_ray_xy(4)
    dirs = O.unproject_ndc_depth1_dirs(cams, xy)
    centers = O.camera_centers(cams)
    pts = centers[:, None, :] + 1.7 * dirs
    back = O.transform_points_ndc(cams, pts)
    assert torch.allclose(back, xy[None].expand_as(back), atol=1e-5)
    assert torch.allclose(dirs.norm(dim=-1), torch.ones(5, 16), atol=1e-6)


def test_lookat_cameras_see_the_origin_at_centre():
    cams = O.lookat_cameras(8, seed=0)
    xy = O.transform_points_ndc(cams, torch.zeros(9, 1, 3))
    assert xy.abs().max() < 1e-5
    assert torch.allclose(O.camera_centers(cams).norm(dim=-1), torch.full((9,), 1.5), atol=1e-5)


def test_patch_grid_and_depths():
    xy = O.patch_ray_xy(2)
    assert torch.allclose(xy, torch.tensor([[0.5, 0.5], [-0.5, 0.5], [0.5, -0.5], [-0.5, -0.5]]))
    depths, deltas = O.raymarcher_depths(24, 2.0)
    assert depths.shape == (24,) and math.isclose(float(depths[0]), 1 / 24, rel_tol=1e-6)
    assert torch.allclose(deltas, torch.full((24,), 1 / 12))
