"""CPU oracle for frame assembly / output formats (scope row f-3).

TEST INFRASTRUCTURE ONLY (see oracle/nerf_oracle.py header).

Restates, in numpy, what the reference does between the render outputs and the bytes it writes:

  unflatten_reshape      models/nerf_downX_model.py:410-416   [(h1 w1) (s1 s2) c] -> [(h1 s1) (w1 s2) c]
  depth2im               utils/visualizer.py:164-176          nan_to_num, (x - near) / max(far - near, 1e-8),
                                                              (255 x).astype(uint8), cv2 COLORMAP_JET, / 255.
  calculate_vis          models/nerf_downX_model.py:418-450   [pred | gt | depth] concatenated along the width
  _save_image            utils/visualizer.py:40-60            ((img - lo) / (hi - lo) * 255.).astype(uint8)
                                                              (the RGB->BGR swap + PNG encode happen after this point)
  _save_matrix           utils/visualizer.py:93-99            np.nan_to_num(depth matrix) -> .npz

Third-party arithmetic: numpy's float32 -> uint8 ``astype`` (C cast: truncate toward zero through a 32-bit
integer, low byte kept, so out-of-range values WRAP -- depths below `near` are common: depth = sum w z with
opacity < 1) and OpenCV's COLORMAP_JET table (opencv-python, unpinned in the reference's requirements.txt)."""
from __future__ import annotations

import numpy as np


def astype_u8(x: np.ndarray) -> np.ndarray:
    """numpy's float -> uint8 cast on x86-64: cvttss2si (out of range / NaN -> INT32_MIN), keep the low byte."""
    x = np.asarray(x, dtype=np.float32)
    ok = np.isfinite(x) & (np.abs(x) < 2147483648.0)
    t = np.trunc(np.where(ok, x, 0.0)).astype(np.int64)
    return (t & 0xFF).astype(np.uint8)


def jet_lut() -> np.ndarray:
    """[256,3] uint8 in the channel order cv2.applyColorMap returns."""
    import cv2
    return cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_JET)[0]


def unflatten_reshape(x: np.ndarray, H: int, W: int, s: int) -> np.ndarray:
    """models/nerf_downX_model.py:410-416.  x: [H*W, C] (or [H*W]) in LR-pixel-major / sub-pixel-minor order."""
    h1, w1 = H // s, W // s
    x = np.asarray(x).reshape(h1, w1, s, s, -1)
    return x.transpose(0, 2, 1, 3, 4).reshape(H, W, -1)


def depth2im(depth_hw: np.ndarray, near: float, far: float) -> np.ndarray:
    """utils/visualizer.py:164-176 -> float32 [H,W,3]."""
    x = np.nan_to_num(np.asarray(depth_hw, dtype=np.float32))
    near, far = np.float32(near), np.float32(far)
    x = (x - near) / max(far - near, 1e-8)
    x = astype_u8(255 * x)
    return (jet_lut()[x] / 255.).astype(np.float32)


def save_image_u8(img: np.ndarray, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    """utils/visualizer.py:51-55 up to (not including) the RGB->BGR swap."""
    img = np.asarray(img, dtype=np.float32)
    return astype_u8((img - lo) / (hi - lo) * 255.)


def assemble_frame(rgb: np.ndarray, depth: np.ndarray, H: int, W: int, s: int, near: float, far: float, gt=None):
    """calculate_vis + _save_image for one (rgb, depth) pair: returns (uint8 [H, W*(2|3), 3], depth matrix [H,W] fp32).
    rgb [H*W,3], depth [H*W], gt [H*W,3] or None, all in the grouped row order (s == 1: plain raster)."""
    img = unflatten_reshape(rgb, H, W, s).astype(np.float32)
    d = unflatten_reshape(depth, H, W, s)[..., 0].astype(np.float32)
    panels = [img]
    if gt is not None:
        panels.append(unflatten_reshape(gt, H, W, s).astype(np.float32))
    panels.append(depth2im(d, near, far))
    pred = np.concatenate(panels, axis=1)
    return save_image_u8(pred), np.nan_to_num(d)
