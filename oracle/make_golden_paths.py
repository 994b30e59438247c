"""Generate tests/golden/pose_paths.npz from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Pins nerf_sr_b200/paths.py (scope row f-4) against the reference's own
create_spiral_poses / create_spheric_poses / average_poses / center_poses (data/llff_downX_dataset.py:20-160)."""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerf_sr_b200 import paths as P    # noqa: E402
from oracle import ref_shim            # noqa: E402


def main():
    ref_shim._install_stubs()
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    if "kornia" not in sys.modules:            # imported by the dataset module, unused by the path functions
        sys.modules["kornia"] = types.ModuleType("kornia")
    from data import llff_downX_dataset as D
    g = np.random.default_rng(3)
    out = {}
    radii, focus = np.array([0.31, 0.22, 0.08]), 3.7
    ref = D.create_spiral_poses(radii, focus, 24)
    mine = P.spiral_poses(radii, focus, 24)
    assert np.allclose(ref, mine, rtol=0, atol=1e-15), np.abs(ref - mine).max()
    out["spiral_in"], out["spiral"] = np.array([*radii, focus, 24.0]), ref
    ref = D.create_spheric_poses(1.37, 20)
    mine = P.spheric_poses(1.37, 20)
    assert np.allclose(ref, mine, rtol=0, atol=1e-15), np.abs(ref - mine).max()
    out["spheric_in"], out["spheric"] = np.array([1.37, 20.0]), ref
    q = np.stack([np.linalg.qr(np.eye(3) + 0.2 * g.standard_normal((3, 3)))[0] for _ in range(7)])
    poses = np.concatenate([q, g.standard_normal((7, 3, 1))], -1)
    ref_c, ref_avg = D.center_poses(poses)
    mine_c, mine_avg = P.center_poses(poses)
    assert np.allclose(ref_c, mine_c, rtol=0, atol=1e-14) and np.allclose(ref_avg, mine_avg, rtol=0, atol=1e-15)
    out["poses"], out["centered"], out["avg"] = poses, ref_c, ref_avg
    path = os.path.join(ROOT, "tests", "golden", "pose_paths.npz")
    np.savez_compressed(path, **out)
    print("pose_paths ok", os.path.getsize(path))


if __name__ == "__main__":
    main()
