"""Generate tests/golden/scene_llff.npz and scene_blender.npz from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Run:  python oracle/make_golden_scenes.py

Pins nerf_sr_b200/scenes.py (scope row f-4, second half: the COLMAP / Blender-JSON loaders) against the reference's
own dataset classes.  A tiny synthetic capture is written to a temporary directory in the real on-disk formats
(COLMAP binary sparse model + PNG images; transforms_*.json + RGBA PNGs), the reference's LLFFDownXDataset /
BlenderDownXDataset (data/llff_downX_dataset.py, data/blender_downX_dataset.py) read it, and their buffers are stored
next to the raw file bytes, so the tests can re-materialise the capture anywhere and compare loader against loader.
The script asserts the pin (bit-equal poses / bounds / targets) before writing."""
from __future__ import annotations

import io
import json
import os
import struct
import sys
import tempfile
import types
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerf_sr_b200 import scenes as S   # noqa: E402
from oracle import nerf_oracle as O    # noqa: E402
from oracle import ref_shim            # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


# ---- a synthetic capture in COLMAP's binary format ----------------------------------------------------
def rotmat_to_qvec(R: np.ndarray) -> np.ndarray:
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    x = (R[2, 1] - R[1, 2]) / (4 * w)
    y = (R[0, 2] - R[2, 0]) / (4 * w)
    z = (R[1, 0] - R[0, 1]) / (4 * w)
    return np.array([w, x, y, z])


def synth_colmap(g: np.random.Generator, n_img: int, n_pts: int, W: int, H: int):
    """World-to-camera poses on a small forward-facing rig looking down +z (COLMAP convention: right-down-front),
    a point cloud in front of it, per-point visibility tracks."""
    cams, images, points = [], [], []
    cams.append(dict(id=1, model=2, w=W, h=H, params=[1.1 * W, W / 2, H / 2, 0.01]))        # SIMPLE_RADIAL f, cx, cy, k
    names = [f"IMG_{i:03d}.png" for i in g.permutation(n_img)]                                # file order != name order
    for k in range(n_img):
        ang = 0.08 * g.standard_normal(3)
        cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
        R = (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
             @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))
        centre = np.array([0.6 * g.standard_normal(), 0.4 * g.standard_normal(), 0.1 * g.standard_normal()])
        t = -R @ centre
        n2d = int(g.integers(0, 5))
        images.append(dict(id=k + 1, q=rotmat_to_qvec(R), t=t, cam=1, name=names[k],
                           pts2d=[(float(g.uniform(0, W)), float(g.uniform(0, H)), int(g.integers(-1, n_pts))) for _ in range(n2d)]))
    for p in range(n_pts):
        xyz = np.array([2.5 * g.standard_normal(), 1.8 * g.standard_normal(), g.uniform(4.0, 30.0)])
        seen = sorted(g.choice(n_img, size=int(g.integers(2, n_img + 1)), replace=False) + 1)
        points.append(dict(id=10 + 3 * p, xyz=xyz, rgb=g.integers(0, 256, 3), err=float(g.uniform(0.1, 2.0)),
                           track=[(int(i), int(g.integers(0, 100))) for i in seen]))
    return cams, images, points


def write_colmap(dirname: str, cams, images, points) -> None:
    os.makedirs(dirname, exist_ok=True)
    with open(os.path.join(dirname, "cameras.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(cams)))
        for c in cams:
            f.write(struct.pack("<iiQQ", c["id"], c["model"], c["w"], c["h"]))
            f.write(struct.pack(f"<{len(c['params'])}d", *c["params"]))
    with open(os.path.join(dirname, "images.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(images)))
        for im in images:
            f.write(struct.pack("<i4d3di", im["id"], *im["q"], *im["t"], im["cam"]))
            f.write(im["name"].encode() + b"\x00")
            f.write(struct.pack("<Q", len(im["pts2d"])))
            for x, y, pid in im["pts2d"]:
                f.write(struct.pack("<ddq", x, y, pid))
    with open(os.path.join(dirname, "points3D.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(points)))
        for p in points:
            f.write(struct.pack("<Q3d3BdQ", p["id"], *p["xyz"], *[int(v) for v in p["rgb"]], p["err"], len(p["track"])))
            for i, j in p["track"]:
                f.write(struct.pack("<ii", i, j))


def synth_image(g: np.random.Generator, W: int, H: int, channels: int) -> bytes:
    from PIL import Image
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.stack([127 + 120 * np.sin(xx * g.uniform(0.1, 0.6) + yy * g.uniform(0.1, 0.6) + g.uniform(0, 6)) for _ in range(3)], -1)
    img = np.clip(img + g.normal(0, 6, img.shape), 0, 255).astype(np.uint8)
    if channels == 4:
        alpha = (255 * (((xx - W / 2) ** 2 + (yy - H / 2) ** 2) < (0.42 * W) ** 2)).astype(np.uint8)
        alpha[::5, ::3] = 128
        img = np.concatenate([img, alpha[..., None]], -1)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="PNG")
    return buf.getvalue()


def dataset_opt(**kw) -> Namespace:
    base = dict(dataset_root=None, img_wh=None, downscale=2, ds_method="lanc", use_pixel_centers=True, sisr_path=None,
                spheric_poses=False, val_num=1, unified_dir=False, all_ref=False, include_var=False, use_subset=False,
                subset_num=20, with_ref=False, no_ref_loss=False, reg_patch_len=1, patch_len=32, rand_dir=False, precrop_frac=0.5)
    base.update(kw)
    return Namespace(**base)


def import_reference_datasets():
    ref_shim._install_stubs()
    if ref_shim.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    if "kornia" not in sys.modules:
        sys.modules["kornia"] = types.ModuleType("kornia")
    from data.blender_downX_dataset import BlenderDownXDataset
    from data.llff_downX_dataset import LLFFDownXDataset
    return LLFFDownXDataset, BlenderDownXDataset


def build_llff(LLFF) -> dict:
    g = np.random.default_rng(20)
    W0, H0, n_img = 40, 30, 6
    img_wh = (24, 18)
    arrays = {}
    with tempfile.TemporaryDirectory() as root:
        cams, images, points = synth_colmap(g, n_img, 60, W0, H0)
        write_colmap(os.path.join(root, "sparse", "0"), cams, images, points)
        os.makedirs(os.path.join(root, "images"))
        os.makedirs(os.path.join(root, "sisr"))
        for im in images:
            data = synth_image(g, W0, H0, 3)
            open(os.path.join(root, "images", im["name"]), "wb").write(data)
            arrays["file_images/" + im["name"]] = np.frombuffer(data, np.uint8)
            data = synth_image(g, *img_wh, 3)
            open(os.path.join(root, "sisr", im["name"]), "wb").write(data)
            arrays["file_sisr/" + im["name"]] = np.frombuffer(data, np.uint8)
        for fn in ("cameras.bin", "images.bin", "points3D.bin"):
            arrays["file_sparse/0/" + fn] = np.frombuffer(open(os.path.join(root, "sparse", "0", fn), "rb").read(), np.uint8)

        meta = dict(img_wh=list(img_wh), torch=torch.__version__, cases=[])
        for tag, kw in (("ndc_lanc_s2", dict(downscale=2, ds_method="lanc")),
                        ("ndc_avg_s3_sr", dict(downscale=3, ds_method="avg", sisr_path=os.path.join(root, "sisr"))),
                        ("spheric_lanc_s2", dict(downscale=2, ds_method="lanc", spheric_poses=True)),
                        ("ndc_unified_dir_s2", dict(downscale=2, ds_method="lanc", unified_dir=True)),
                        ("spheric_corner_pixels_s3", dict(downscale=3, ds_method="avg", spheric_poses=True, use_pixel_centers=False))):
            opt = dataset_opt(dataset_root=root, img_wh=img_wh, **kw)
            ref = LLFF(opt, "train")
            s = opt.downscale
            mine = S.load_llff_scene(root, img_wh, spheric_poses=opt.spheric_poses, sisr_path=opt.sisr_path,
                                     use_pixel_centers=opt.use_pixel_centers, unified_dir=opt.unified_dir)
            # ---- the pin: scene-level quantities bit-equal ----
            assert mine.focal == ref.focal and mine.val_idx == int(np.argmin(np.linalg.norm(ref.poses[..., 3], axis=1)))
            assert np.array_equal(mine.poses, ref.poses), np.abs(mine.poses - ref.poses).max()
            assert np.array_equal(mine.bounds, ref.bounds)
            assert [os.path.basename(p) for p in mine.image_paths] == [os.path.basename(p) for p in ref.image_paths]
            lr, hr, sr = [], [], []
            for i in mine.train_indices():
                a, b = S.load_image_targets(mine.image_paths[i], img_wh, s, opt.ds_method)
                lr.append(a), hr.append(b)
                if opt.sisr_path:
                    sr.append(S.load_sr_target(mine.sr_image_paths[i], img_wh, s))
            assert np.array_equal(np.concatenate(lr), ref.all_rgbs.numpy()), tag
            assert np.array_equal(np.concatenate(hr), ref.all_rgbs_ori.numpy()), tag
            if opt.sisr_path:
                assert np.array_equal(np.concatenate(sr), ref.all_rgbs_sr.numpy()), tag
            # rays: the oracle's restatement of the dataset path (pinned in raygen.npz) on my poses
            rays = torch.cat([O.build_frame_rays(torch.from_numpy(mine.poses[i]).float(), img_wh[1], img_wh[0], mine.focal, s,
                                                 mine.near, mine.far, mine.ndc, mine.use_pixel_centers,
                                                 mine.unified_dir).view(-1, s * s, 8)
                              for i in mine.train_indices()], 0)
            assert torch.equal(rays, ref.all_rays), (tag, float((rays - ref.all_rays).abs().max()))
            test = LLFF(opt, "test")
            assert np.allclose(mine.test_poses("test"), test.poses_test, rtol=0, atol=1e-14)
            val = LLFF(opt, "val")
            vs = val[0]
            vlr, vhr = S.load_image_targets(mine.image_paths[mine.val_idx], img_wh, s, "avg")
            assert np.array_equal(vlr, vs["rgbs"].numpy()) and np.array_equal(vhr, vs["rgbs_ori"].numpy())
            arrays[f"{tag}/poses"], arrays[f"{tag}/bounds"] = ref.poses, ref.bounds
            arrays[f"{tag}/all_rays"], arrays[f"{tag}/all_rgbs"] = ref.all_rays.numpy(), ref.all_rgbs.numpy()
            arrays[f"{tag}/all_rgbs_ori"] = ref.all_rgbs_ori.numpy()
            if opt.sisr_path:
                arrays[f"{tag}/all_rgbs_sr"] = ref.all_rgbs_sr.numpy()
            arrays[f"{tag}/poses_test"] = np.asarray(test.poses_test)
            arrays[f"{tag}/val_rays"], arrays[f"{tag}/val_rgbs"] = vs["rays"].numpy(), vs["rgbs"].numpy()
            arrays[f"{tag}/val_rgbs_ori"] = vs["rgbs_ori"].numpy()
            meta["cases"].append(dict(tag=tag, downscale=s, ds_method=opt.ds_method, spheric_poses=opt.spheric_poses,
                                      sisr=bool(opt.sisr_path), use_pixel_centers=opt.use_pixel_centers,
                                      unified_dir=opt.unified_dir, focal=float(ref.focal), val_idx=int(mine.val_idx),
                                      near=mine.near, far=mine.far))
    arrays["meta_json"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    return arrays


def build_blender(Blender) -> dict:
    g = np.random.default_rng(21)
    arrays = {}
    img_wh = (16, 16)
    with tempfile.TemporaryDirectory() as root:
        meta = dict(img_wh=list(img_wh), torch=torch.__version__, cases=[])
        for split, n in (("train", 3), ("test", 2)):
            os.makedirs(os.path.join(root, split))
            frames = []
            for i in range(n):
                q, _ = np.linalg.qr(np.eye(3) + 0.5 * g.standard_normal((3, 3)))
                c2w = np.eye(4)
                c2w[:3, :3], c2w[:3, 3] = q, 4.0 * q[:, 2]
                frames.append(dict(file_path=f"./{split}/r_{i}", rotation=0.0126, transform_matrix=c2w.tolist()))
                data = synth_image(g, 20, 20, 4)
                open(os.path.join(root, split, f"r_{i}.png"), "wb").write(data)
                arrays[f"file_{split}/r_{i}.png"] = np.frombuffer(data, np.uint8)
            text = json.dumps(dict(camera_angle_x=0.6911112070083618, frames=frames), indent=1).encode()
            open(os.path.join(root, f"transforms_{split}.json"), "wb").write(text)
            arrays[f"file_transforms_{split}.json"] = np.frombuffer(text, np.uint8)
        for tag, kw in (("lanc_s2", dict(downscale=2, ds_method="lanc")), ("avg_s4", dict(downscale=4, ds_method="avg"))):
            opt = dataset_opt(dataset_root=root, img_wh=img_wh, **kw)
            s = opt.downscale
            ref = Blender(opt, "train")
            mine = S.load_blender_scene(root, "train", img_wh)
            assert mine.focal == ref.focal and np.array_equal(mine.poses, np.stack(ref.poses))
            lr, hr = zip(*[S.load_image_targets(p, img_wh, s, opt.ds_method, rgba=True) for p in mine.image_paths])
            assert np.array_equal(np.concatenate(lr), ref.all_rgbs.numpy()), tag
            assert np.array_equal(np.concatenate(hr), ref.all_rgbs_ori.numpy()), tag
            rays = torch.cat([O.build_frame_rays(torch.from_numpy(p).float(), img_wh[1], img_wh[0], mine.focal, s, 2.0, 6.0,
                                                 False).view(-1, s * s, 8) for p in mine.poses], 0)
            assert torch.equal(rays, ref.all_rays), (tag, float((rays - ref.all_rays).abs().max()))
            test = Blender(opt, "test")
            ts = test[1]
            tm = S.load_blender_scene(root, "test", img_wh)
            tlr, thr = S.load_image_targets(tm.image_paths[1], img_wh, s, opt.ds_method, rgba=True)
            assert np.array_equal(tlr, ts["rgbs"].numpy()) and np.array_equal(thr, ts["rgbs_ori"].numpy())
            arrays[f"{tag}/poses"] = np.stack(ref.poses)
            arrays[f"{tag}/all_rays"], arrays[f"{tag}/all_rgbs"] = ref.all_rays.numpy(), ref.all_rgbs.numpy()
            arrays[f"{tag}/all_rgbs_ori"] = ref.all_rgbs_ori.numpy()
            arrays[f"{tag}/test1_rays"], arrays[f"{tag}/test1_rgbs"] = ts["rays"].numpy(), ts["rgbs"].numpy()
            arrays[f"{tag}/test1_rgbs_ori"] = ts["rgbs_ori"].numpy()
            meta["cases"].append(dict(tag=tag, downscale=s, ds_method=opt.ds_method, focal=float(ref.focal)))
        # the 'train_crop' split (--precrop_frac 0.5): central window of targets and rays
        opt = dataset_opt(dataset_root=root, img_wh=img_wh, downscale=2, ds_method="lanc", precrop_frac=0.5)
        ref = Blender(opt, "train_crop")
        mine = S.load_blender_scene(root, "train_crop", img_wh)
        keep = S.center_crop_lr_indices(img_wh, 2, 0.5)
        lr, hr, rays = [], [], []
        for p_img, pose in zip(mine.image_paths, mine.poses):
            a, b = S.load_image_targets(p_img, img_wh, 2, "lanc", rgba=True)
            lr.append(a[keep]), hr.append(b[keep])
            rays.append(O.build_frame_rays(torch.from_numpy(pose).float(), img_wh[1], img_wh[0], mine.focal, 2, 2.0, 6.0,
                                           False).view(-1, 4, 8)[torch.from_numpy(keep)])
        assert np.array_equal(np.concatenate(lr), ref.all_rgbs.numpy()) and np.array_equal(np.concatenate(hr), ref.all_rgbs_ori.numpy())
        assert torch.equal(torch.cat(rays, 0), ref.all_rays)
        arrays["crop_lanc_s2/all_rays"], arrays["crop_lanc_s2/all_rgbs"] = ref.all_rays.numpy(), ref.all_rgbs.numpy()
        arrays["crop_lanc_s2/all_rgbs_ori"] = ref.all_rgbs_ori.numpy()
        meta["crop"] = dict(tag="crop_lanc_s2", downscale=2, ds_method="lanc", precrop_frac=0.5, n_keep=int(keep.size))
    arrays["meta_json"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    return arrays


def main() -> None:
    LLFF, Blender = import_reference_datasets()
    for name, arrays in (("scene_llff", build_llff(LLFF)), ("scene_blender", build_blender(Blender))):
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: pinned to the reference dataset class, {os.path.getsize(path) / 1e3:.1f} KB, {len(arrays)} arrays")


if __name__ == "__main__":
    main()
