"""Output files of a test sweep (scope row f-3, second half): what the reference's ``save_visuals`` leaves on disk.

``Renderer.render_test_pose`` / ``render_path`` end where the reference hands pixels to an encoder: uint8 RGB frames
and fp32 depth matrices (``nsr_assemble_frame``).  This module writes them in the reference's formats and names
(models/nerf_downX_model.py:626-669, utils/visualizer.py:40-60,100-105):

    <i>-coarse.png  <i>-fine.png  <i>-coarse-ori.png  <i>-fine-ori.png      [pred | depth] panels, LR and HR
    <i>-coarse-depth.npz  <i>-fine-depth.npz  <i>-*-depth-ori.npz          np.savez(mat) -> key 'arr_0'; the
                                                                            ``*-fine-depth-ori.npz`` files are what warp.py:100-112 reads
    coarse.gif  fine.gif  coarse-ori.gif  fine-ori.gif                      30 fps sweeps (optional)

PNG is lossless, so the encoder is irrelevant to parity: what must match is the decoded pixel array, and the reference's
``cv2.cvtColor(RGB2BGR)`` + ``cv2.imwrite`` (which expects BGR) stores exactly the RGB array it was given.  ``encode_png``
is a dependency-free writer (zlib from the standard library); tests decode its output with cv2 and PIL.  The reference's
GIFs go through imageio's palette quantiser (``palettesize=256``); ``write_gif`` uses PIL's, so GIF pixels are not a parity
surface.  ``_save_matrix``'s debugging side file (``<name>test-depth.png``) is not written."""
from __future__ import annotations

import os
import struct
import zlib
from typing import Dict, Iterable, Mapping, Optional, Sequence

import numpy as np


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def encode_png(rgb8: np.ndarray, level: int = 6) -> bytes:
    """uint8 [H, W, 3] (RGB) or [H, W] (grey) -> PNG file bytes (8 bit, no interlace, 'up' row filter)."""
    a = np.ascontiguousarray(rgb8)
    if a.dtype != np.uint8 or a.ndim not in (2, 3) or (a.ndim == 3 and a.shape[2] not in (1, 3)):
        raise ValueError(f"encode_png expects uint8 [H,W,3] or [H,W], got {a.dtype} {a.shape}")
    h, w = a.shape[:2]
    if h == 0 or w == 0:
        raise ValueError("encode_png: empty image")
    colour = 2 if (a.ndim == 3 and a.shape[2] == 3) else 0
    rows = a.reshape(h, -1)
    # filter type 2 (Up): row - previous row, mod 256; the first row's predecessor is all zeros
    up = rows.copy()
    up[1:] = rows[1:] - rows[:-1]
    raw = np.concatenate([np.full((h, 1), 2, dtype=np.uint8), up], axis=1).tobytes()
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, colour, 0, 0, 0))
            + _chunk(b"IDAT", zlib.compress(raw, level)) + _chunk(b"IEND", b""))


def _host(x) -> np.ndarray:
    """A numpy view of a CPU / CUDA tensor or an array (frames arrive as device or pinned host tensors)."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def write_png(path: str, rgb8) -> None:
    """The file ``_save_image`` writes for an HWC uint8 RGB panel (utils/visualizer.py:40-60)."""
    with open(path, "wb") as fh:
        fh.write(encode_png(_host(rgb8)))


def write_depth_npz(path: str, depth_mat) -> None:
    """The file ``_save_matrix`` writes: ``np.savez(path, np.nan_to_num(mat))`` (utils/visualizer.py:100-105).  The
    matrices ``nsr_assemble_frame`` returns already have NaNs zeroed on the device."""
    np.savez(path, np.nan_to_num(_host(depth_mat)))


def write_gif(path: str, frames: Iterable, fps: int = 30) -> None:
    """An animated GIF of uint8 RGB frames (the reference's ``_save_gif``, utils/visualizer.py:63-80, via PIL)."""
    from PIL import Image
    imgs = [Image.fromarray(_host(f)) for f in frames]
    if not imgs:
        raise ValueError("write_gif: no frames")
    imgs[0].save(path, save_all=True, append_images=imgs[1:], duration=int(round(1000 / fps)), loop=0)


FRAME_FILES = {"coarse_pred": "{i}-coarse.png", "fine_pred": "{i}-fine.png", "coarse_pred_ori": "{i}-coarse-ori.png",
               "fine_pred_ori": "{i}-fine-ori.png", "coarse_depth_mat": "{i}-coarse-depth.npz",
               "fine_depth_mat": "{i}-fine-depth.npz", "coarse_depth_mat_ori": "{i}-coarse-depth-ori.npz",
               "fine_depth_mat_ori": "{i}-fine-depth-ori.npz"}


def save_test_frame(out_dir: str, index: int, frame: Mapping[str, object], keys: Optional[Sequence[str]] = None) -> Dict[str, str]:
    """Write the files the reference's ``test()`` + ``save_visuals`` produce for pose ``index`` from one
    ``Renderer.render_test_pose`` / ``render_path`` result.  Returns {key: path} of what was written."""
    os.makedirs(out_dir, exist_ok=True)
    written = {}
    for k in (keys if keys is not None else [k for k in FRAME_FILES if k in frame]):
        if k not in FRAME_FILES:
            raise KeyError(f"{k!r} is not a test-sweep output (known: {sorted(FRAME_FILES)})")
        path = os.path.join(out_dir, FRAME_FILES[k].format(i=index))
        (write_depth_npz if path.endswith(".npz") else write_png)(path, frame[k])
        written[k] = path
    return written


def save_test_sweep(out_dir: str, frames: Iterable[Mapping[str, object]], gif_keys: Sequence[str] = ()) -> int:
    """Consume a ``Renderer.render_path`` generator: per-pose files as ``save_test_frame``, plus one GIF per key in
    ``gif_keys`` (e.g. ('fine_pred', 'fine_pred_ori') -> fine.gif, fine-ori.gif).  Returns the number of poses."""
    gifs = {k: [] for k in gif_keys}
    n = 0
    for n, frame in enumerate(frames, 1):
        save_test_frame(out_dir, n - 1, frame)
        for k in gifs:
            gifs[k].append(_host(frame[k]).copy())        # render_path reuses its pinned slots
    for k, imgs in gifs.items():
        name = FRAME_FILES[k].format(i="").lstrip("-").replace(".png", ".gif")
        write_gif(os.path.join(out_dir, name), imgs)
    return n
