"""GPU: scene loaders -> device buffers -> the render / training path (scope row f-4 end to end, through the C ABI).
The ray buffers the reference's dataset classes build on the CPU (tests/golden/scene_*.npz) against the ones
``Scene.train_buffers`` generates on the device; one training iteration and one test-sweep frame from a loaded scene."""
import os

import numpy as np
import pytest
import torch

from conftest import materialize_scene
from nerf_sr_b200 import scenes as S
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _renderer(cfg, seeds=(4, 17)):
    from nerf_sr_b200 import Renderer
    r = Renderer(cfg, torch.device(DEV), precision="bf16x3")
    pc, pf = O.make_mlp_params(cfg, seeds[0]), O.make_mlp_params(cfg, seeds[1])
    r.load_state_dict(0, pc)
    r.load_state_dict(1, pf)
    return r, pc, pf


def test_llff_train_buffers_match_reference_dataset(tmp_path):
    z, meta = materialize_scene("scene_llff", str(tmp_path))
    r, _, _ = _renderer(O.RenderConfig())
    for case in meta["cases"]:
        tag, s = case["tag"], case["downscale"]
        sc = S.load_llff_scene(str(tmp_path), meta["img_wh"], spheric_poses=case["spheric_poses"],
                               sisr_path=os.path.join(str(tmp_path), "sisr") if case["sisr"] else None,
                               use_pixel_centers=case["use_pixel_centers"], unified_dir=case["unified_dir"])
        buf = sc.train_buffers(r, s, case["ds_method"], with_sr=case["sisr"])
        ref = torch.from_numpy(z[f"{tag}/all_rays"])
        assert buf["rays"].shape == ref.shape and buf["rays"].is_cuda
        # same tolerance as the raygen golden test: the rotation's multiply-adds round differently from torch's CPU matmul
        assert torch.allclose(buf["rays"].cpu(), ref, rtol=1e-5, atol=1e-5), float((buf["rays"].cpu() - ref).abs().max())
        assert torch.equal(buf["rays"][..., 6:].cpu(), ref[..., 6:])                      # near / far columns exact
        assert np.array_equal(buf["rgbs"].cpu().numpy(), z[f"{tag}/all_rgbs"])
        assert np.array_equal(buf["rgbs_ori"].cpu().numpy(), z[f"{tag}/all_rgbs_ori"])
        if case["sisr"]:
            assert np.array_equal(buf["rgbs_sr"].cpu().numpy(), z[f"{tag}/all_rgbs_sr"])
        val = sc.val_sample(r, s)
        vref = torch.from_numpy(z[f"{tag}/val_rays"]).reshape(-1, 8)
        assert torch.allclose(val["rays"].cpu(), vref, rtol=1e-5, atol=1e-5)
        assert np.array_equal(val["rgbs"].cpu().numpy(), z[f"{tag}/val_rgbs"])
    r.close()


def test_blender_train_buffers_match_reference_dataset(tmp_path):
    z, meta = materialize_scene("scene_blender", str(tmp_path))
    r, _, _ = _renderer(O.RenderConfig(white_bkgd=True))
    for case in meta["cases"]:
        tag, s = case["tag"], case["downscale"]
        sc = S.load_blender_scene(str(tmp_path), "train", meta["img_wh"])
        buf = sc.train_buffers(r, s, case["ds_method"])
        ref = torch.from_numpy(z[f"{tag}/all_rays"])
        assert torch.allclose(buf["rays"].cpu(), ref, rtol=1e-5, atol=1e-5), float((buf["rays"].cpu() - ref).abs().max())
        assert np.array_equal(buf["rgbs"].cpu().numpy(), z[f"{tag}/all_rgbs"])
        te = S.load_blender_scene(str(tmp_path), "test", meta["img_wh"])
        rays = te.frame_rays(r, te.poses[1], s)
        assert torch.allclose(rays.cpu(), torch.from_numpy(z[f"{tag}/test1_rays"]).reshape(-1, 8), rtol=1e-5, atol=1e-5)
    crop = meta["crop"]                                    # the 'train_crop' split: central window, gathered on the device
    sc = S.load_blender_scene(str(tmp_path), "train_crop", meta["img_wh"])
    buf = sc.train_buffers(r, crop["downscale"], crop["ds_method"], precrop_frac=crop["precrop_frac"])
    ref = torch.from_numpy(z["crop_lanc_s2/all_rays"])
    assert buf["rays"].shape == ref.shape and torch.allclose(buf["rays"].cpu(), ref, rtol=1e-5, atol=1e-5)
    assert np.array_equal(buf["rgbs"].cpu().numpy(), z["crop_lanc_s2/all_rgbs"])
    assert np.array_equal(buf["rgbs_ori"].cpu().numpy(), z["crop_lanc_s2/all_rgbs_ori"])
    assert S.append_viewdir(buf["rays"].reshape(-1, 8)).shape[1] == 11
    r.close()


def test_scene_to_training_step_and_test_frame(tmp_path):
    """files -> Scene -> device buffers -> Trainer.optimize_parameters (all loss terms) -> a test-sweep frame -> files."""
    from nerf_sr_b200 import Trainer, frames as F
    _, meta = materialize_scene("scene_llff", str(tmp_path))
    cfg = O.RenderConfig(noise_std=1.0)
    r, pc, pf = _renderer(cfg, (21, 8))
    sc = S.load_llff_scene(str(tmp_path), meta["img_wh"], sisr_path=os.path.join(str(tmp_path), "sisr"))
    s = 2
    buf = sc.train_buffers(r, s, "lanc", with_sr=True)
    n = buf["rays"].shape[0]
    assert n == 5 * 12 * 9
    tr = Trainer(r, pc, pf, downscale=s, lambda_coarse_var=0.01, lambda_fine_var=0.01, lambda_coarse_depth_var=0.01,
                 lambda_fine_depth_var=0.01)
    g = torch.Generator(device=DEV).manual_seed(0)
    fixed = S.take_batch(buf, torch.arange(0, n, 7, device=DEV))               # a fixed probe batch, eval-mode sampling

    def probe():
        tr.forward_backward(fixed["rays"], fixed["rgbs"], None, target_sr=fixed["rgbs_sr"], far=sc.far)
        return float(tr.last_terms[:, 5].sum())
    first = probe()
    for it in range(8):
        idx = torch.randperm(n, device=DEV, generator=g)[:64]
        b = S.take_batch(buf, idx)
        assert b["rays"].shape == (64 * s * s, 8) and b["rgbs"].shape == (64, 3) and b["rgbs_sr"].shape == (64 * s * s, 3)
        tr.optimize_parameters(b["rays"], b["rgbs"], tr.draw_rng(b["rays"].shape[0], g), target_sr=b["rgbs_sr"], far=sc.far)
        assert np.isfinite(float(tr.last_terms[:, 5].sum()))
    last = probe()
    assert np.isfinite(first) and last < first, (first, last)
    # --with_ref: the reference view's sub-pixel groups feed the second forward of the iteration
    ref = sc.ref_buffers(r, s)
    assert ref["ref_rays"].shape == (12 * 9, s * s, 8) and ref["ref_rgbs"].shape == (12 * 9, s * s, 3)
    rb = S.take_batch(ref, torch.arange(0, 10, device=DEV))
    b = S.take_batch(buf, torch.arange(0, 64, device=DEV))
    tr.optimize_parameters(b["rays"], b["rgbs"], tr.draw_rng(b["rays"].shape[0], g), target_sr=b["rgbs_sr"], far=sc.far,
                           ref_rays=rb["ref_rays"], ref_rgbs=rb["ref_rgbs"], ref_rng=tr.draw_rng(rb["ref_rays"].shape[0], g))
    assert torch.isfinite(tr.last_ref_terms).all() and float(tr.last_ref_terms.sum()) > 0
    w, h = sc.img_wh
    frames = list(sc.render_sweep(r, s, n_poses=2))
    assert len(frames) == 2 and tuple(frames[0]["fine_pred_ori"].shape) == (h, 2 * w, 3)
    out = os.path.join(str(tmp_path), "results")
    F.save_test_frame(out, 0, frames[1])
    assert np.load(os.path.join(out, "0-fine-depth-ori.npz"))["arr_0"].shape == (h, w)
    r.close()
