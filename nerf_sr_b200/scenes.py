"""Scene loaders: the data formats on the input side of the render path (scope row f-4, second half).

Host-side mirror of what the reference's dataset classes read from disk and how they turn it into the
render path's inputs.  Only the file parsing and the per-scene pose algebra run on the host (a scene is a
few dozen 3x4 matrices); everything per ray runs on the device -- the ray buffers the reference builds on
the CPU with get_ray_directions / get_rays / get_ndc_rays / einops.rearrange (one [H*W, 8] tensor per image)
are produced by ``nsr_generate_rays`` straight into HBM, 48 bytes of pose up per image.

    read_cameras_binary / read_images_binary / read_points3d_binary
                                  COLMAP's binary sparse model (utils/colmap.py:108-135, :168-200, :230-257; the format is
                                  COLMAP's published ``src/base/reconstruction.cc`` layout, little endian)
    load_llff_scene               LLFFDownXDataset.read_meta steps 1-3 (data/llff_downX_dataset.py:197-262): focal
                                  rescale, w2c -> c2w, per-image depth bounds from the sparse points, axis flip,
                                  pose centring, val-image choice, scene rescale, near / far, NDC or spheric
    load_blender_scene            BlenderDownXDataset.read_meta (data/blender_downX_dataset.py:70-90): transforms_*.json
    Scene.test_poses              the test-sweep paths (data/llff_downX_dataset.py:373-385)
    load_image_targets            image -> (LR target, grouped HR target) (:311-330; blender :104-121,139-147)
    Scene.train_buffers           the reference's all_rays / all_rgbs / all_rgbs_ori (/ all_rgbs_sr) buffers, on the device
    Scene.frame_rays              one frame's rays (val / test / test_train samples, :473-494)
    center_crop_lr_indices        the Blender 'train_crop' window (data/blender_downX_dataset.py:123-150)
    append_viewdir                the vanilla datasets' 11-column rows (data/llff_dataset.py:337-341)

Image decoding and resampling are PIL's (the reference's own third-party dependency for exactly this:
``Image.open(..).convert('RGB').resize(.., Image.LANCZOS)``); nothing here touches the oracle.  Results are pinned
to the reference's dataset classes in tests/golden/scene_*.npz (oracle/make_golden_scenes.py).

Not mirrored (the reference path stays): ``--rand_dir`` (per-dataset random sub-pixel jitter from numpy's global RNG) and
the 'gan' / 'reg_patch' random-patch sampling modes."""
from __future__ import annotations

import json
import os
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import paths

# COLMAP camera models: id -> (name, number of parameters)
CAMERA_MODELS = {0: ("SIMPLE_PINHOLE", 3), 1: ("PINHOLE", 4), 2: ("SIMPLE_RADIAL", 4), 3: ("RADIAL", 5), 4: ("OPENCV", 8),
                 5: ("OPENCV_FISHEYE", 8), 6: ("FULL_OPENCV", 12), 7: ("FOV", 5), 8: ("SIMPLE_RADIAL_FISHEYE", 4),
                 9: ("RADIAL_FISHEYE", 5), 10: ("THIN_PRISM_FISHEYE", 12)}


@dataclass
class ColmapCamera:
    id: int
    model: str
    width: int
    height: int
    params: np.ndarray


@dataclass
class ColmapImage:
    id: int
    qvec: np.ndarray        # (w, x, y, z)
    tvec: np.ndarray
    camera_id: int
    name: str

    def rotmat(self) -> np.ndarray:
        return qvec_to_rotmat(self.qvec)


def qvec_to_rotmat(q: np.ndarray) -> np.ndarray:
    """Unit quaternion (w, x, y, z) -> 3x3 rotation, the expression of utils/colmap.py:272-282."""
    w, x, y, z = q
    return np.array([[1 - 2 * y ** 2 - 2 * z ** 2, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
                     [2 * x * y + 2 * w * z, 1 - 2 * x ** 2 - 2 * z ** 2, 2 * y * z - 2 * w * x],
                     [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x ** 2 - 2 * y ** 2]])


class _Cursor:
    """Little-endian reader over a bytes object (whole file in memory: sparse models are a few MB)."""

    def __init__(self, data: bytes, what: str):
        self.d, self.o, self.what = data, 0, what

    def take(self, fmt: str):
        size = struct.calcsize("<" + fmt)
        if self.o + size > len(self.d):
            raise ValueError(f"{self.what}: truncated file (need {size} bytes at offset {self.o}, have {len(self.d) - self.o})")
        v = struct.unpack_from("<" + fmt, self.d, self.o)
        self.o += size
        return v

    def array(self, dtype, count: int) -> np.ndarray:
        dt = np.dtype(dtype).newbyteorder("<")
        size = dt.itemsize * count
        if self.o + size > len(self.d):
            raise ValueError(f"{self.what}: truncated file (need {size} bytes at offset {self.o}, have {len(self.d) - self.o})")
        a = np.frombuffer(self.d, dtype=dt, count=count, offset=self.o)
        self.o += size
        return a

    def cstring(self) -> str:
        end = self.d.find(b"\x00", self.o)
        if end < 0:
            raise ValueError(f"{self.what}: unterminated string at offset {self.o}")
        s = self.d[self.o:end].decode("utf-8")
        self.o = end + 1
        return s


def _read(path: str) -> _Cursor:
    with open(path, "rb") as fh:
        return _Cursor(fh.read(), path)


def read_cameras_binary(path: str) -> Dict[int, ColmapCamera]:
    """cameras.bin: u64 count, then per camera {i32 id, i32 model, u64 width, u64 height, f64 params[n(model)]}."""
    c = _read(path)
    out: Dict[int, ColmapCamera] = {}
    for _ in range(c.take("Q")[0]):
        cam_id, model_id, width, height = c.take("iiQQ")
        if model_id not in CAMERA_MODELS:
            raise ValueError(f"{path}: unknown camera model id {model_id}")
        name, n_params = CAMERA_MODELS[model_id]
        out[cam_id] = ColmapCamera(cam_id, name, int(width), int(height), c.array(np.float64, n_params).copy())
    return out


def read_images_binary(path: str) -> List[ColmapImage]:
    """images.bin: u64 count, then per image {i32 id, f64 q[4], f64 t[3], i32 camera, name\\0, u64 n2d,
    n2d x {f64 x, f64 y, i64 point3d}}.  Returned in FILE order (the reference iterates its dict in insertion
    order, data/llff_downX_dataset.py:211-216); the 2D observations are skipped."""
    c = _read(path)
    out: List[ColmapImage] = []
    for _ in range(c.take("Q")[0]):
        image_id = c.take("i")[0]
        qvec = c.array(np.float64, 4).copy()
        tvec = c.array(np.float64, 3).copy()
        camera_id = c.take("i")[0]
        name = c.cstring()
        n2d = c.take("Q")[0]
        c.array(np.uint8, 24 * n2d)
        out.append(ColmapImage(image_id, qvec, tvec, camera_id, name))
    return out


def read_points3d_binary(path: str) -> Tuple[np.ndarray, List[np.ndarray]]:
    """points3D.bin: u64 count, then per point {u64 id, f64 xyz[3], u8 rgb[3], f64 error, u64 track,
    track x {i32 image_id, i32 point2d_idx}}.  Returns (xyz [P,3] in file order, per-point image-id arrays)."""
    c = _read(path)
    n = c.take("Q")[0]
    xyz = np.zeros((n, 3))
    tracks: List[np.ndarray] = []
    for i in range(n):
        c.take("Q")
        xyz[i] = c.array(np.float64, 3)
        c.take("BBBd")
        t = c.take("Q")[0]
        tracks.append(c.array(np.int32, 2 * t).reshape(t, 2)[:, 0].copy())
    return xyz, tracks


@dataclass
class Scene:
    """What the render path needs to know about a captured scene."""
    kind: str                          # 'llff' | 'blender'
    root: str
    img_wh: Tuple[int, int]            # (W, H) of the HR rays (opt.img_wh)
    focal: float
    poses: np.ndarray                  # [F,3,4] float64 camera-to-world, right-up-back
    image_paths: List[str]
    bounds: np.ndarray                 # llff: [F,2] per-image (near, far) depth percentiles; blender: [2]
    ndc: bool                          # forward-facing scene rendered in NDC (near 0, far 1)
    near: float                        # the values written into ray columns 6, 7
    far: float
    white_back: bool
    val_idx: Optional[int] = None      # llff: the image closest to the centre (held out from training)
    spheric: bool = False
    rgba: bool = False                 # blender PNGs carry alpha, blended onto white
    use_pixel_centers: bool = True     # --use_pixel_centers (options/base_options.py:59)
    unified_dir: bool = False          # --unified_dir (llff): one view direction per LR pixel
    sr_image_paths: List[str] = field(default_factory=list)

    # ---- poses ------------------------------------------------------------------------------------
    def train_indices(self, include_val: bool = False) -> List[int]:
        """Images whose rays go into the training buffers (data/llff_downX_dataset.py:298-300)."""
        return [i for i in range(len(self.image_paths)) if include_val or i != self.val_idx]

    def test_poses(self, split: str = "test", n_poses: int = 120) -> np.ndarray:
        """Poses of a test sweep (data/llff_downX_dataset.py:373-385): the training poses for ``*_train`` splits, a
        spiral through the 90th percentile of the camera offsets for forward-facing scenes (focus depth 3.5), a circle
        of radius 1.1 x nearest bound for spheric ones; for Blender scenes the split's own poses."""
        if self.kind == "blender" or split.endswith("train"):
            return self.poses
        if not self.spheric:
            radii = np.percentile(np.abs(self.poses[..., 3]), 90, axis=0)
            return paths.spiral_poses(radii, 3.5, n_poses)
        return paths.spheric_poses(1.1 * self.bounds.min(), n_poses)

    # ---- rays (device) ----------------------------------------------------------------------------
    def frame_rays(self, renderer, pose: np.ndarray, s: int):
        """[H*W, 8] device rays of one pose in LR-pixel-major / sub-pixel-minor row order: the 'rays' entry of a
        val / test sample flattened (data/llff_downX_dataset.py:473-494, data/blender_downX_dataset.py:207-215)."""
        w, h = self.img_wh
        if w % s or h % s:
            raise ValueError(f"img_wh {self.img_wh} is not divisible by downscale {s}")
        return renderer.generate_rays(np.asarray(pose, dtype=np.float32), h, w, self.focal, s, self.ndc, self.near, self.far,
                                      use_pixel_centers=self.use_pixel_centers, unified_dir=self.unified_dir)

    def train_buffers(self, renderer, s: int, ds_method: str = "lanc", include_val: bool = False, with_sr: bool = False,
                      precrop_frac: Optional[float] = None):
        """The reference's training buffers on the renderer's device: ``rays`` [n, s*s, 8] (all_rays), ``rgbs`` [n, 3]
        (all_rgbs, the LR targets), ``rgbs_ori`` [n, s*s, 3] (all_rgbs_ori) and, ``with_sr``, ``rgbs_sr`` [n, s*s, 3]
        (all_rgbs_sr), n = images x (H/s) x (W/s).  Rays are generated on the device per pose; targets are decoded on the
        host (PIL) and uploaded once.  ``precrop_frac``: the Blender 'train_crop' split (``--precrop_frac``,
        data/blender_downX_dataset.py:123-150): only the LR pixels of the central window are kept."""
        import torch
        w, h = self.img_wh
        keep = None if precrop_frac is None else center_crop_lr_indices(self.img_wh, s, precrop_frac)
        keep_dev = None if keep is None else torch.from_numpy(keep).to(renderer.device)
        rays, rgbs, rgbs_ori, rgbs_sr = [], [], [], []
        for i in self.train_indices(include_val):
            r = self.frame_rays(renderer, self.poses[i], s).view(-1, s * s, 8)
            lr, hr = load_image_targets(self.image_paths[i], self.img_wh, s, ds_method, rgba=self.rgba)
            sr = load_sr_target(self.sr_image_paths[i], self.img_wh, s) if with_sr else None
            if keep is not None:
                r, lr, hr = r[keep_dev], lr[keep], hr[keep]
                sr = None if sr is None else sr[keep]
            rays.append(r)
            rgbs.append(torch.from_numpy(lr))
            rgbs_ori.append(torch.from_numpy(hr))
            if with_sr:
                rgbs_sr.append(torch.from_numpy(sr))
        dev = renderer.device
        out = {"rays": torch.cat(rays, 0), "rgbs": torch.cat(rgbs, 0).to(dev), "rgbs_ori": torch.cat(rgbs_ori, 0).to(dev)}
        if with_sr:
            out["rgbs_sr"] = torch.cat(rgbs_sr, 0).to(dev)
        return out


    def ref_buffers(self, renderer, s: int, ref_idx: int = 0):
        """``--with_ref`` (data/llff_downX_dataset.py:288-293, 331-332, 357-358): the reference view's rays and HR colours,
        ``ref_rays`` [n_lr, s*s, 8] / ``ref_rgbs`` [n_lr, s*s, 3], from which every training sample draws one LR pixel's
        sub-pixel group (``take_batch`` with a random index gives the flattened [B*s*s, .] batch ``Trainer`` takes)."""
        import torch
        _, hr = load_image_targets(self.image_paths[ref_idx], self.img_wh, s, "lanc", rgba=self.rgba)
        return {"ref_rays": self.frame_rays(renderer, self.poses[ref_idx], s).view(-1, s * s, 8),
                "ref_rgbs": torch.from_numpy(hr).to(renderer.device)}

    def val_sample(self, renderer, s: int, index: Optional[int] = None):
        """The reference's 'val' / 'test_train' sample of image ``index`` (default: the held-out val image): device
        ``rays`` [H*W, 8], ``rgbs`` [H*W/s^2, 3] (always the s x s mean of the HR image, data/llff_downX_dataset.py:499-507)
        and ``rgbs_ori`` [H*W, 3] in ray row order."""
        import torch
        i = self.val_idx if index is None else index
        if i is None:
            raise ValueError("this scene has no held-out image; pass index")
        lr, hr = load_image_targets(self.image_paths[i], self.img_wh, s, "avg" if self.kind == "llff" else "lanc", rgba=self.rgba)
        dev = renderer.device
        return {"rays": self.frame_rays(renderer, self.poses[i], s), "rgbs": torch.from_numpy(lr).to(dev),
                "rgbs_ori": torch.from_numpy(hr).reshape(-1, 3).to(dev)}

    def render_sweep(self, renderer, s: int, split: str = "test", n_poses: int = 120, **kw):
        """``Renderer.render_path`` over this scene's test poses with its intrinsics: a generator of per-pose host
        frames (see nerf_sr_b200.frames.save_test_sweep for the files the reference writes from them)."""
        w, h = self.img_wh
        return renderer.render_path(self.test_poses(split, n_poses), h, w, self.focal, s, self.ndc, self.near, self.far, **kw)


def center_crop_lr_indices(img_wh: Sequence[int], s: int, precrop_frac: float = 0.5) -> np.ndarray:
    """Row-major indices of the LR pixels inside the Blender 'train_crop' window (data/blender_downX_dataset.py:123-131):
    the reference crops the LR image to rows ``H_lr//2 +- int(H_lr//2 * frac)`` (columns likewise) and the HR image / rays to
    ``H//2 +- int(H//2 * frac)`` before grouping.  The two windows describe the same pixels only when the HR one is exactly s x
    the LR one (true for the reference's sizes); anything else would pair rays with the wrong targets, so it is refused.
    (The reference's own reshape additionally only works for frac = 0.5, where the window is half the frame.)"""
    w, h = int(img_wh[0]), int(img_wh[1])
    if w % s or h % s:
        raise ValueError(f"img_wh {tuple(img_wh)} is not divisible by downscale {s}")
    w_lr, h_lr = w // s, h // s
    dh_lr, dw_lr = int(h_lr // 2 * precrop_frac), int(w_lr // 2 * precrop_frac)
    dh, dw = int(h // 2 * precrop_frac), int(w // 2 * precrop_frac)
    r0, c0 = h_lr // 2 - dh_lr, w_lr // 2 - dw_lr
    if (h // 2 - dh, w // 2 - dw, 2 * dh, 2 * dw) != (s * r0, s * c0, s * 2 * dh_lr, s * 2 * dw_lr):
        raise ValueError(f"precrop window of the {w}x{h} rays is not {s}x the window of the {w_lr}x{h_lr} targets")
    if dh_lr < 1 or dw_lr < 1:
        raise ValueError("empty precrop window")
    rows = np.arange(r0, r0 + 2 * dh_lr)[:, None]
    cols = np.arange(c0, c0 + 2 * dw_lr)[None, :]
    return (rows * w_lr + cols).reshape(-1).astype(np.int64)


def append_viewdir(rays):
    """[N, 8] -> [N, 11]: the vanilla datasets' row layout (o, d, near, far, viewdir = d; data/llff_dataset.py:337-341,
    data/blender_dataset.py), read by ``NeRFModel.forward_rays`` at columns 8:11 (models/nerf_model.py:213)."""
    import torch
    return torch.cat([rays, rays[:, 3:6]], 1)


def take_batch(buffers, index):
    """One training batch from ``Scene.train_buffers``: what the reference's DataLoader + ``set_input`` hand to
    ``forward`` (models/nerf_downX_model.py:238-248: every [B, s*s, C] entry is flattened to [B*s*s, C]).
    ``index``: LR-pixel indices (a device LongTensor, e.g. a slice of torch.randperm)."""
    out = {}
    for k, v in buffers.items():
        b = v[index]
        out[k] = b.reshape(-1, b.shape[-1]) if b.ndim == 3 else b
    return out


# ---- scenes ---------------------------------------------------------------------------------------------
def load_llff_scene(root: str, img_wh: Sequence[int], spheric_poses: bool = False, use_subset: bool = False,
                    subset_num: int = 20, sisr_path: Optional[str] = None, use_pixel_centers: bool = True,
                    unified_dir: bool = False) -> Scene:
    """A COLMAP-reconstructed real scene (``<root>/sparse/0/*.bin`` + ``<root>/images``), normalised exactly like
    LLFFDownXDataset.read_meta (data/llff_downX_dataset.py:197-262)."""
    cams = read_cameras_binary(os.path.join(root, "sparse/0/cameras.bin"))
    if 1 not in cams:
        raise ValueError("cameras.bin has no camera with id 1 (the reference reads camdata[1])")
    focal = cams[1].params[0] * img_wh[0] / cams[1].width                     # :200-203
    images = read_images_binary(os.path.join(root, "sparse/0/images.bin"))
    names = [im.name for im in images]
    perm = np.argsort(names)                                                   # :207
    image_paths = [os.path.join(root, "images", n) for n in sorted(names)]     # :209-210
    w2c = np.zeros((len(images), 4, 4))
    for k, im in enumerate(images):                                            # :211-218
        w2c[k, :3, :3] = im.rotmat()
        w2c[k, :3, 3] = im.tvec
        w2c[k, 3, 3] = 1.0
    poses = np.linalg.inv(w2c)[:, :3]                                          # :219 camera-to-world, file order

    xyz, tracks = read_points3d_binary(os.path.join(root, "sparse/0/points3D.bin"))
    n_img, n_pts = len(poses), len(xyz)
    visible = np.zeros((n_img, n_pts), dtype=bool)
    for i, ids in enumerate(tracks):                                           # :226-229: row = image_id - 1
        rows = ids - 1
        if rows.size and (rows.min() < -n_img or rows.max() >= n_img):
            raise ValueError(f"points3D.bin: image id out of range 1..{n_img} (the reference indexes visibilities[id-1])")
        visible[rows, i] = True
    pts_world = xyz.T[None]                                                    # [1,3,P]
    depths = ((pts_world - poses[..., 3:4]) * poses[..., 2:3]).sum(1)          # :232 depth along each camera's front axis
    bounds = np.zeros((n_img, 2))
    for i in range(n_img):                                                     # :233-236
        zs = depths[i][visible[i]]
        bounds[i] = [np.percentile(zs, 0.1), np.percentile(zs, 99.9)]
    poses, bounds = poses[perm], bounds[perm]                                  # :238-239 name order
    poses = np.concatenate([poses[..., 0:1], -poses[..., 1:3], poses[..., 3:4]], -1)   # :243 right-down-front -> right-up-back
    poses, _ = paths.center_poses(poses)                                       # :244
    val_idx = int(np.argmin(np.linalg.norm(poses[..., 3], axis=1)))            # :245-246
    scale = bounds.min() * 0.75                                                # :253-257 nearest depth at 1 / 0.75
    bounds = bounds / scale
    poses[..., 3] /= scale

    sr_paths: List[str] = []
    if sisr_path is not None:                                                  # :258-263
        sr_paths = [os.path.join(sisr_path, f) for f in sorted(os.listdir(sisr_path))
                    if f.endswith("JPG") or f.endswith("jpg") or f.endswith("png")]
        if use_subset:
            sr_paths = sr_paths[:subset_num]
    if use_subset:                                                             # :265-267 (val_idx is chosen before the cut)
        poses, image_paths = poses[:subset_num], image_paths[:subset_num]
    if spheric_poses:                                                          # :340-344
        near = float(bounds.min())
        far = float(min(8 * near, bounds.max()))
    else:                                                                      # :333-338 NDC: near plane 1.0 -> (0, 1)
        near, far = 0.0, 1.0
    return Scene("llff", root, (int(img_wh[0]), int(img_wh[1])), float(focal), poses, image_paths, bounds,
                 ndc=not spheric_poses, near=near, far=far, white_back=False, val_idx=val_idx, spheric=spheric_poses,
                 sr_image_paths=sr_paths, use_pixel_centers=use_pixel_centers, unified_dir=unified_dir)


def load_blender_scene(root: str, split: str, img_wh: Sequence[int], use_pixel_centers: bool = True) -> Scene:
    """A synthetic NeRF scene (``transforms_<split>.json`` + RGBA PNGs), BlenderDownXDataset.read_meta
    (data/blender_downX_dataset.py:70-90, :101-107)."""
    if img_wh[0] != img_wh[1]:
        raise ValueError("image width must equal image height")
    split_path = "train" if split == "train_crop" else split
    with open(os.path.join(root, f"transforms_{split_path}.json")) as fh:
        meta = json.load(fh)
    focal = 0.5 * 800 / np.tan(0.5 * meta["camera_angle_x"])                   # :76 focal at W = 800
    focal *= img_wh[0] / 800                                                   # :79
    poses = np.stack([np.array(f["transform_matrix"])[:3, :4] for f in meta["frames"]], 0)
    image_paths = [os.path.join(root, f"{f['file_path']}.png") for f in meta["frames"]]
    return Scene("blender", root, (int(img_wh[0]), int(img_wh[1])), float(focal), poses, image_paths, np.array([2.0, 6.0]),
                 ndc=False, near=2.0, far=6.0, white_back=True, rgba=True, use_pixel_centers=use_pixel_centers)


# ---- images -> targets -----------------------------------------------------------------------------------
def _to_unit_float(img) -> np.ndarray:
    """torchvision ToTensor for 8-bit images, channels last: uint8 / 255 as float32, [H,W,C]."""
    a = np.asarray(img, dtype=np.uint8)
    if a.ndim == 2:
        a = a[..., None]
    return a.astype(np.float32) / np.float32(255.0)


def group_subpixels(img: np.ndarray, s: int) -> np.ndarray:
    """'(h s1) (w s2) c -> (h w) (s1 s2) c' (data/llff_downX_dataset.py:328-329): HR raster -> [n_lr, s*s, C]."""
    H, W, Cn = img.shape
    return np.ascontiguousarray(img.reshape(H // s, s, W // s, s, Cn).transpose(0, 2, 1, 3, 4).reshape((H // s) * (W // s), s * s, Cn))


def _blend_white(rgba: np.ndarray) -> np.ndarray:
    """rgb * a + (1 - a) (data/blender_downX_dataset.py:119-121)."""
    return rgba[..., :3] * rgba[..., 3:4] + (np.float32(1.0) - rgba[..., 3:4])


def load_image_targets(path: str, img_wh: Sequence[int], s: int, ds_method: str = "lanc", rgba: bool = False):
    """(lr [H/s * W/s, 3], hr [H/s * W/s, s*s, 3]) float32 supervision of one image: open, LANCZOS-resize to
    ``img_wh``, then the LR target by a second LANCZOS resize ('lanc') or an s x s mean ('avg')
    (data/llff_downX_dataset.py:311-330; alpha-blended onto white for Blender, data/blender_downX_dataset.py:104-121)."""
    from PIL import Image
    w, h = int(img_wh[0]), int(img_wh[1])
    img = Image.open(path)
    if not rgba:
        img = img.convert("RGB")
    img = img.resize((w, h), Image.LANCZOS)
    hr = _to_unit_float(img)
    if rgba and hr.shape[-1] != 4:
        raise ValueError(f"{path}: expected an RGBA image")
    if ds_method == "lanc":
        lr = _to_unit_float(img.resize((w // s, h // s), Image.LANCZOS))
    elif ds_method == "avg":
        # F.avg_pool2d(img, s): windows over the top-left (H//s * s, W//s * s) region, fp32 sum / s^2
        hh, ww = (h // s) * s, (w // s) * s
        win = hr[:hh, :ww].reshape(h // s, s, w // s, s, hr.shape[-1]).transpose(0, 2, 1, 3, 4).reshape(h // s, w // s, s * s, -1)
        acc = np.zeros(win.shape[:2] + win.shape[3:], dtype=np.float32)
        for k in range(s * s):
            acc += win[:, :, k]
        lr = acc / np.float32(s * s)
    else:
        raise ValueError("Downscale option not found")
    if rgba:
        hr, lr = _blend_white(hr), _blend_white(lr)
    return np.ascontiguousarray(lr.reshape(-1, 3)), group_subpixels(hr, s)


def load_sr_target(path: str, img_wh: Sequence[int], s: int) -> np.ndarray:
    """[H/s * W/s, s*s, 3]: a SISR-upscaled image used as HR supervision (``--sisr_path``,
    data/llff_downX_dataset.py:301-309); it must already have the HR size."""
    from PIL import Image
    img = Image.open(path).convert("RGB")
    if img.size[0] != img_wh[0] or img.size[1] != img_wh[1]:
        raise ValueError("sr image sizes mismatch")
    return group_subpixels(_to_unit_float(img), s)
