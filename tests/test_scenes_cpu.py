"""CPU: the scene loaders (nerf_sr_b200/scenes.py, scope row f-4) against the buffers the reference's own dataset
classes produced from the same files (tests/golden/scene_*.npz, oracle/make_golden_scenes.py): COLMAP binary parsing,
pose normalisation, depth bounds, val-image choice, LR / HR / SISR targets, test-sweep poses, Blender JSON; plus the
parsers' behaviour on damaged files."""
import os
import struct

import numpy as np
import pytest
import torch

from conftest import materialize_scene
from nerf_sr_b200 import scenes as S
from oracle import nerf_oracle as O
from oracle import ref_shim


def _llff_case(root, meta, case):
    return S.load_llff_scene(root, meta["img_wh"], spheric_poses=case["spheric_poses"],
                             sisr_path=os.path.join(root, "sisr") if case["sisr"] else None,
                             use_pixel_centers=case["use_pixel_centers"], unified_dir=case["unified_dir"])


def test_llff_scene_matches_reference_dataset(tmp_path):
    z, meta = materialize_scene("scene_llff", str(tmp_path))
    assert len(meta["cases"]) == 5
    for case in meta["cases"]:
        tag, s = case["tag"], case["downscale"]
        sc = _llff_case(str(tmp_path), meta, case)
        assert sc.focal == case["focal"] and sc.val_idx == case["val_idx"]
        assert (sc.near, sc.far) == (case["near"], case["far"]) and sc.ndc == (not case["spheric_poses"])
        # same numpy expressions on the same bytes: equal to the last bit in the build container; LAPACK's inverse may
        # round differently on another host, hence 1e-12
        assert np.allclose(sc.poses, z[f"{tag}/poses"], rtol=0, atol=1e-12)
        assert np.allclose(sc.bounds, z[f"{tag}/bounds"], rtol=1e-12, atol=0)
        assert [os.path.basename(p) for p in sc.image_paths] == sorted(os.path.basename(p) for p in sc.image_paths)
        lr, hr, sr = [], [], []
        for i in sc.train_indices():
            a, b = S.load_image_targets(sc.image_paths[i], sc.img_wh, s, case["ds_method"])
            lr.append(a), hr.append(b)
            if case["sisr"]:
                sr.append(S.load_sr_target(sc.sr_image_paths[i], sc.img_wh, s))
        assert sc.val_idx not in sc.train_indices() and len(sc.train_indices()) == len(sc.image_paths) - 1
        assert np.array_equal(np.concatenate(lr), z[f"{tag}/all_rgbs"])            # PIL decode + LANCZOS / s x s mean
        assert np.array_equal(np.concatenate(hr), z[f"{tag}/all_rgbs_ori"])
        if case["sisr"]:
            assert np.array_equal(np.concatenate(sr), z[f"{tag}/all_rgbs_sr"])
        assert np.allclose(sc.test_poses("test"), z[f"{tag}/poses_test"], rtol=0, atol=1e-12)
        assert sc.test_poses("test_train") is sc.poses
        vlr, vhr = S.load_image_targets(sc.image_paths[sc.val_idx], sc.img_wh, s, "avg")   # val samples always use the mean
        assert np.array_equal(vlr, z[f"{tag}/val_rgbs"]) and np.array_equal(vhr, z[f"{tag}/val_rgbs_ori"])
        # the ray buffers: the oracle's dataset-path restatement on the loader's poses == the reference's all_rays
        w, h = sc.img_wh
        rays = torch.cat([O.build_frame_rays(torch.from_numpy(sc.poses[i]).float(), h, w, sc.focal, s, sc.near, sc.far,
                                             sc.ndc, sc.use_pixel_centers, sc.unified_dir).view(-1, s * s, 8) for i in sc.train_indices()], 0)
        assert torch.allclose(rays, torch.from_numpy(z[f"{tag}/all_rays"]), rtol=1e-6, atol=1e-6)


def test_blender_scene_matches_reference_dataset(tmp_path):
    z, meta = materialize_scene("scene_blender", str(tmp_path))
    for case in meta["cases"]:
        tag, s = case["tag"], case["downscale"]
        sc = S.load_blender_scene(str(tmp_path), "train", meta["img_wh"])
        assert sc.focal == case["focal"] and (sc.near, sc.far, sc.ndc, sc.white_back) == (2.0, 6.0, False, True)
        assert np.array_equal(sc.poses, z[f"{tag}/poses"])
        lr, hr = zip(*[S.load_image_targets(p, sc.img_wh, s, case["ds_method"], rgba=True) for p in sc.image_paths])
        assert np.array_equal(np.concatenate(lr), z[f"{tag}/all_rgbs"])            # incl. the alpha blend onto white
        assert np.array_equal(np.concatenate(hr), z[f"{tag}/all_rgbs_ori"])
        te = S.load_blender_scene(str(tmp_path), "test", meta["img_wh"])
        tlr, thr = S.load_image_targets(te.image_paths[1], te.img_wh, s, case["ds_method"], rgba=True)
        assert np.array_equal(tlr, z[f"{tag}/test1_rgbs"]) and np.array_equal(thr, z[f"{tag}/test1_rgbs_ori"])
        assert te.test_poses("test") is te.poses and len(te.poses) == 2
    with pytest.raises(ValueError):
        S.load_blender_scene(str(tmp_path), "train", (12, 8))
    # the 'train_crop' split: central window of targets and (oracle-built) rays == the reference's cropped buffers
    crop = meta["crop"]
    sc = S.load_blender_scene(str(tmp_path), "train_crop", meta["img_wh"])
    keep = S.center_crop_lr_indices(sc.img_wh, crop["downscale"], crop["precrop_frac"])
    assert keep.size == crop["n_keep"] == 16 and keep[0] == 2 * 8 + 2 and keep[-1] == 5 * 8 + 5
    lr, hr, rays = [], [], []
    for p_img, pose in zip(sc.image_paths, sc.poses):
        a, b = S.load_image_targets(p_img, sc.img_wh, 2, "lanc", rgba=True)
        lr.append(a[keep]), hr.append(b[keep])
        rays.append(O.build_frame_rays(torch.from_numpy(pose).float(), 16, 16, sc.focal, 2, 2.0, 6.0, False).view(-1, 4, 8)[torch.from_numpy(keep)])
    assert np.array_equal(np.concatenate(lr), z["crop_lanc_s2/all_rgbs"]) and np.array_equal(np.concatenate(hr), z["crop_lanc_s2/all_rgbs_ori"])
    assert torch.allclose(torch.cat(rays, 0), torch.from_numpy(z["crop_lanc_s2/all_rays"]), rtol=1e-6, atol=1e-6)


def test_center_crop_window_rules():
    assert S.center_crop_lr_indices((400, 400), 2, 0.5).size == 100 * 100            # the reference's Blender size
    assert S.center_crop_lr_indices((800, 800), 4, 0.5).size == 100 * 100
    with pytest.raises(ValueError, match="not 2x the window"):
        S.center_crop_lr_indices((12, 12), 2, 0.5)       # LR window 2x2 but HR window 6x6: the reference's buffers would disagree
    with pytest.raises(ValueError, match="divisible"):
        S.center_crop_lr_indices((15, 15), 2, 0.5)
    with pytest.raises(ValueError):
        S.center_crop_lr_indices((16, 16), 2, 0.0)
    rays = torch.arange(24, dtype=torch.float32).view(3, 8)
    r11 = S.append_viewdir(rays)
    assert r11.shape == (3, 11) and torch.equal(r11[:, :8], rays) and torch.equal(r11[:, 8:], rays[:, 3:6])


def test_colmap_parsers_reject_damaged_files(tmp_path):
    materialize_scene("scene_llff", str(tmp_path))
    sparse = os.path.join(str(tmp_path), "sparse", "0")
    cams = S.read_cameras_binary(os.path.join(sparse, "cameras.bin"))
    assert cams[1].model == "SIMPLE_RADIAL" and (cams[1].width, cams[1].height) == (40, 30) and cams[1].params.shape == (4,)
    imgs = S.read_images_binary(os.path.join(sparse, "images.bin"))
    assert [im.id for im in imgs] == list(range(1, 7)) and all(im.name.endswith(".png") for im in imgs)
    R = imgs[0].rotmat()
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.linalg.det(R) == pytest.approx(1.0)
    xyz, tracks = S.read_points3d_binary(os.path.join(sparse, "points3D.bin"))
    assert xyz.shape == (60, 3) and len(tracks) == 60 and all(2 <= len(t) <= 6 for t in tracks)
    for fn, reader in (("cameras.bin", S.read_cameras_binary), ("images.bin", S.read_images_binary),
                       ("points3D.bin", S.read_points3d_binary)):
        data = open(os.path.join(sparse, fn), "rb").read()
        cut = os.path.join(str(tmp_path), "cut_" + fn)
        open(cut, "wb").write(data[: len(data) - 5])
        with pytest.raises(ValueError, match="truncated"):
            reader(cut)
    bad = os.path.join(str(tmp_path), "bad_model.bin")
    open(bad, "wb").write(struct.pack("<QiiQQ", 1, 1, 99, 4, 4))
    with pytest.raises(ValueError, match="unknown camera model"):
        S.read_cameras_binary(bad)
    empty = os.path.join(str(tmp_path), "empty.bin")
    open(empty, "wb").write(struct.pack("<Q", 0))
    assert S.read_cameras_binary(empty) == {} and S.read_images_binary(empty) == []
    assert S.read_points3d_binary(empty)[0].shape == (0, 3)
    with pytest.raises(ValueError, match="Downscale option"):
        S.load_image_targets(os.path.join(str(tmp_path), "images", imgs[0].name), (24, 18), 2, "bicubic")
    with pytest.raises(ValueError, match="mismatch"):
        S.load_sr_target(os.path.join(str(tmp_path), "images", imgs[0].name), (24, 18), 2)


def test_group_subpixels_is_the_dataset_rearrange():
    import einops
    x = np.arange(6 * 8 * 3, dtype=np.float32).reshape(6, 8, 3)
    for s in (1, 2):
        want = einops.rearrange(x, "(h s1) (w s2) c -> (h w) (s1 s2) c", s1=s, s2=s)
        assert np.array_equal(S.group_subpixels(x, s), want)


def test_take_batch_flattens_like_set_input():
    buf = {"rays": torch.arange(5 * 4 * 8, dtype=torch.float32).view(5, 4, 8), "rgbs": torch.rand(5, 3),
           "rgbs_ori": torch.rand(5, 4, 3)}
    b = S.take_batch(buf, torch.tensor([3, 1]))
    assert b["rays"].shape == (8, 8) and b["rgbs"].shape == (2, 3) and b["rgbs_ori"].shape == (8, 3)
    assert torch.equal(b["rays"][:4], buf["rays"][3]) and torch.equal(b["rgbs"][1], buf["rgbs"][1])


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree only exists in the build container")
def test_llff_loader_equals_live_reference_dataset(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import make_golden_scenes as M
    LLFF, _ = M.import_reference_datasets()
    _, meta = materialize_scene("scene_llff", str(tmp_path))
    opt = M.dataset_opt(dataset_root=str(tmp_path), img_wh=tuple(meta["img_wh"]), downscale=2, ds_method="avg")
    ref = LLFF(opt, "train")
    sc = S.load_llff_scene(str(tmp_path), meta["img_wh"])
    assert np.array_equal(sc.poses, ref.poses) and np.array_equal(sc.bounds, ref.bounds) and sc.focal == ref.focal
    lr = np.concatenate([S.load_image_targets(sc.image_paths[i], sc.img_wh, 2, "avg")[0] for i in sc.train_indices()])
    assert np.array_equal(lr, ref.all_rgbs.numpy())


def test_colmap_reader_roundtrips_a_written_model(tmp_path):
    """Every field the loaders use survives write -> read for several camera models, name orders and track lengths."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import make_golden_scenes as M
    for seed, n_img, n_pts, model in ((1, 3, 5, 0), (2, 9, 40, 1), (3, 4, 0, 4), (4, 2, 17, 10)):
        g = np.random.default_rng(seed)
        cams, images, points = M.synth_colmap(g, n_img, n_pts, 64, 48)
        n_params = S.CAMERA_MODELS[model][1]
        cams[0].update(model=model, params=list(g.random(n_params)))
        d = os.path.join(str(tmp_path), f"m{seed}")
        M.write_colmap(d, cams, images, points)
        got_c = S.read_cameras_binary(os.path.join(d, "cameras.bin"))
        assert list(got_c) == [1] and got_c[1].model == S.CAMERA_MODELS[model][0] and np.array_equal(got_c[1].params, cams[0]["params"])
        got_i = S.read_images_binary(os.path.join(d, "images.bin"))
        assert [im.name for im in got_i] == [im["name"] for im in images]
        for a, b in zip(got_i, images):
            assert a.id == b["id"] and a.camera_id == 1 and np.array_equal(a.qvec, b["q"]) and np.array_equal(a.tvec, b["t"])
        xyz, tracks = S.read_points3d_binary(os.path.join(d, "points3D.bin"))
        assert xyz.shape == (n_pts, 3)
        for k, p in enumerate(points):
            assert np.array_equal(xyz[k], p["xyz"]) and list(tracks[k]) == [i for i, _ in p["track"]]


def test_qvec_to_rotmat_is_a_rotation_for_unit_quaternions():
    g = np.random.default_rng(0)
    for _ in range(20):
        q = g.standard_normal(4)
        q /= np.linalg.norm(q)
        R = S.qvec_to_rotmat(q)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.linalg.det(R) == pytest.approx(1.0, abs=1e-12)
        assert np.allclose(S.qvec_to_rotmat(-q), R, atol=1e-15)          # q and -q are the same rotation
