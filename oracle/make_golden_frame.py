"""Generate tests/golden/frame_assembly.npz from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Pins oracle/frame_oracle.py (scope row f-3) against the reference's own
unflatten_reshape / depth2im / calculate_vis concat / _save_image conversion (models/nerf_downX_model.py:
410-450, utils/visualizer.py:40-60,164-176), bit for bit, then stores inputs and the uint8 frames."""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frame_oracle as F   # noqa: E402
from oracle import ref_shim            # noqa: E402


def main():
    ref_shim._install_stubs()
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    from utils.visualizer import depth2im
    from models.nerf_downX_model import NeRFDownXModel
    out = {}
    g = np.random.default_rng(11)
    for tag, (H, W, s, near, far, with_gt) in {"blender": (12, 16, 2, 2.0, 6.0, True), "llff": (16, 8, 4, 0.0, 1.0, False),
                                               "raster": (5, 7, 1, 2.0, 6.0, True)}.items():
        n = H * W
        rgb = g.random((n, 3), dtype=np.float32)
        rgb[::7] = 1.0
        rgb[3::11] = 0.0
        depth = (near + (far - near) * g.random(n, dtype=np.float32) * 1.3 - 0.4).astype(np.float32)   # below near and beyond far too
        depth[5] = np.nan
        depth[9] = np.inf
        gt = g.random((n, 3), dtype=np.float32) if with_gt else None
        # ---- the reference's code path ----
        fake = types.SimpleNamespace(opt=types.SimpleNamespace(img_wh=(W, H), downscale=s))
        unfl = lambda x: NeRFDownXModel.unflatten_reshape(fake, torch.from_numpy(x))
        near_a, far_a = np.array([near], np.float32)[0], np.array([far], np.float32)[0]
        img = unfl(rgb)
        dim = depth2im(unfl(depth.reshape(n, 1))[..., 0], near_a, far_a)
        parts = [img] + ([unfl(gt)] if with_gt else []) + [dim]
        pred = torch.cat(parts, dim=1).numpy()                       # calculate_vis :425-429
        ref_u8 = ((pred - 0) / (1 - 0) * 255.).astype(np.uint8)       # _save_image :54
        ref_mat = np.nan_to_num(unfl(depth.reshape(n, 1))[..., 0].numpy())      # _save_matrix :96
        # ---- the pin ----
        mine_u8, mine_mat = F.assemble_frame(rgb, depth, H, W, s, near, far, gt)
        assert np.array_equal(ref_u8, mine_u8), tag
        assert np.array_equal(ref_mat, mine_mat), tag
        out[f"{tag}_params"] = np.array([H, W, s, near, far, float(with_gt)], np.float64)
        out[f"{tag}_rgb"], out[f"{tag}_depth"] = rgb, depth
        if with_gt:
            out[f"{tag}_gt"] = gt
        out[f"{tag}_u8"], out[f"{tag}_mat"] = ref_u8, ref_mat
    # float -> uint8 cast table: the numpy cast itself on a sweep of awkward values
    sweep = np.array([-3.5, 300.7, -0.2, 255.9, 256.0, -255.5, -256.5, 1e6, -1e6, 3e9, -3e9, np.nan, np.inf, -np.inf, 0.999, 254.999],
                     np.float32)
    with np.errstate(invalid="ignore"):
        cast = sweep.astype(np.uint8)
    assert np.array_equal(cast, F.astype_u8(sweep))
    out["cast_in"], out["cast_out"] = sweep, cast
    out["jet_lut"] = F.jet_lut()
    path = os.path.join(ROOT, "tests", "golden", "frame_assembly.npz")
    np.savez_compressed(path, **out)
    print("frame_assembly ok", os.path.getsize(path))


if __name__ == "__main__":
    main()
