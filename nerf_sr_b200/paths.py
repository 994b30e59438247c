"""Camera paths for test sweeps (scope row f-4: dataset-side ray construction for whole pose paths).

Host-side mirror of the reference's path generators -- a sweep is a few hundred 3x4 matrices, so this part
is plain numpy; everything per-ray happens on the device (nsr_generate_rays inside Renderer.render_test_pose):

    spiral_poses(radii, focus_depth, n)    data/llff_downX_dataset.py:86-118   forward-facing (NDC) scenes
    spheric_poses(radius, n)               data/llff_downX_dataset.py:121-160  360-degree scenes
    average_pose / center_poses            data/llff_downX_dataset.py:20-83    pose normalisation before NDC

Results are pinned to the reference's own functions in tests/golden/pose_paths.npz (oracle/make_golden_paths.py)."""
from __future__ import annotations

import numpy as np


def _unit(v: np.ndarray) -> np.ndarray:
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _look_frames(z: np.ndarray, up: np.ndarray, centers: np.ndarray) -> np.ndarray:
    """[n,3,4] camera-to-world matrices with columns (x, y, z, centre), x = normalise(up x z), y = z x x."""
    x = _unit(np.cross(np.broadcast_to(up, z.shape), z))
    y = np.cross(z, x)
    return np.stack([x, y, z, centers], axis=-1)


def spiral_poses(radii, focus_depth: float, n_poses: int = 120) -> np.ndarray:
    """Two turns of a spiral around the origin, every camera looking at the plane z = -focus_depth."""
    t = np.linspace(0, 4 * np.pi, n_poses + 1)[:-1]
    centers = np.stack([np.cos(t), -np.sin(t), -np.sin(0.5 * t)], axis=-1) * np.asarray(radii)
    z = _unit(centers - np.array([0, 0, -focus_depth]))
    return _look_frames(z, np.array([0, 1, 0]), centers)


def spheric_poses(radius: float, n_poses: int = 120, phi: float = -np.pi / 5) -> np.ndarray:
    """A circle around the z axis looking 36 degrees downwards: for each angle theta,
    flip @ Ry(theta) @ Rx(phi) @ T(0, -0.9 r, r) (homogeneous 4x4 products), first three rows."""
    th = np.linspace(0, 2 * np.pi, n_poses + 1)[:-1]
    n = th.shape[0]
    trans = np.eye(4)
    trans[1, 3], trans[2, 3] = -0.9 * radius, radius
    rx = np.eye(4)
    rx[1, 1], rx[1, 2], rx[2, 1], rx[2, 2] = np.cos(phi), -np.sin(phi), np.sin(phi), np.cos(phi)
    ry = np.tile(np.eye(4), (n, 1, 1))
    ry[:, 0, 0], ry[:, 0, 2], ry[:, 2, 0], ry[:, 2, 2] = np.cos(th), -np.sin(th), np.sin(th), np.cos(th)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    return (flip @ (ry @ rx @ trans))[:, :3]


def average_pose(poses: np.ndarray) -> np.ndarray:
    """[n,3,4] -> [3,4]: mean centre, mean viewing axis, x = normalise(mean-y x z), y = z x x."""
    center = poses[..., 3].mean(0)
    z = _unit(poses[..., 2].mean(0))
    return _look_frames(z[None], poses[..., 1].mean(0), center[None])[0]


def center_poses(poses: np.ndarray):
    """Express every pose in the frame of the average pose: ([n,3,4] centred poses, [3,4] average pose)."""
    avg = average_pose(poses)
    avg_h = np.eye(4)
    avg_h[:3] = avg
    bottom = np.tile(np.array([0, 0, 0, 1]), (len(poses), 1, 1))
    homo = np.concatenate([poses, bottom], 1)
    return (np.linalg.inv(avg_h) @ homo)[:, :3], avg
