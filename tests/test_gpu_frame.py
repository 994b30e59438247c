"""GPU: nsr_assemble_frame (scope row f-3) bit-exact against the reference-pinned fixture and the oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import frame_oracle as F
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _renderer():
    from nerf_sr_b200 import Renderer
    return Renderer(O.RenderConfig(), torch.device(DEV), precision="bf16x3")


@pytest.mark.parametrize("tag", ["blender", "llff", "raster"])
def test_assemble_frame_matches_reference_golden(tag):
    z = np.load(os.path.join(GOLDEN_DIR, "frame_assembly.npz"))
    H, W, s, near, far, with_gt = z[f"{tag}_params"]
    r = _renderer()
    gt = torch.from_numpy(z[f"{tag}_gt"]).to(DEV) if with_gt else None
    u8, mat = r.assemble_frame(torch.from_numpy(z[f"{tag}_rgb"]).to(DEV), torch.from_numpy(z[f"{tag}_depth"]).to(DEV),
                               int(H), int(W), int(s), float(near), float(far), gt)
    assert np.array_equal(u8.cpu().numpy(), z[f"{tag}_u8"])
    assert np.array_equal(mat.cpu().numpy(), z[f"{tag}_mat"])
    r.close()


def test_assemble_frame_full_size_against_oracle():
    """A rendered 400x400 frame (BASELINE configs[1]) through render -> assemble, bit-exact vs the oracle's assembly."""
    cfg = O.RenderConfig(white_bkgd=True)
    from nerf_sr_b200 import Renderer
    r = Renderer(cfg, torch.device(DEV), precision="bf16x3")
    r.load_state_dict(0, O.make_mlp_params(cfg, 4)); r.load_state_dict(1, O.make_mlp_params(cfg, 17))
    H = W = 400
    c2w = torch.tensor([[1., 0, 0, 0.1], [0, 1, 0, -0.2], [0, 0, 1, 4.0]])
    rays = r.generate_rays(c2w, H, W, 555.0, s=2, near=2.0, far=6.0)
    out = r.forward_rays(rays, want_weights=False)
    u8, mat = r.assemble_frame(out["fine_comp_rgbs"], out["fine_depth"], H, W, 2, 2.0, 6.0)
    with np.errstate(all="ignore"):
        ref_u8, ref_mat = F.assemble_frame(out["fine_comp_rgbs"].cpu().numpy(), out["fine_depth"].cpu().numpy(), H, W, 2, 2.0, 6.0)
    assert np.array_equal(u8.cpu().numpy(), ref_u8) and np.array_equal(mat.cpu().numpy(), ref_mat)
    assert u8.shape == (H, 2 * W, 3)
    from nerf_sr_b200 import NsrError
    with pytest.raises(NsrError):
        r.assemble_frame(out["fine_comp_rgbs"], out["fine_depth"], H, W, 3, 2.0, 6.0)
    r.close()


def test_render_test_pose_and_path_sweep():
    """Scope rows f-4 + f-3 together: a pose path -> per-pose rays on the device -> render -> LR/HR frames, i.e. the
    reference's test sweep (dataset test branch + forward + comp_low_res_output + calculate_vis) with 48 bytes going
    up per frame.  Frames are byte-equal to the oracle's assembly of the same render outputs; the LLFF (NDC) branch
    is checked the same way."""
    from nerf_sr_b200 import Renderer, paths
    cfg = O.RenderConfig(white_bkgd=True)
    r = Renderer(cfg, torch.device(DEV), precision="bf16x3")
    r.load_state_dict(0, O.make_mlp_params(cfg, 4)); r.load_state_dict(1, O.make_mlp_params(cfg, 17))
    H, W, s, focal = 48, 64, 2, 70.0
    poses = paths.spheric_poses(4.0, 5).astype(np.float32)
    frames = []
    for host in r.render_path(poses, H, W, focal, s, ndc=False, near=2.0, far=6.0,
                              keys=("fine_pred", "fine_pred_ori", "fine_depth_mat_ori", "coarse_pred")):
        frames.append({k: v.clone() for k, v in host.items()})
    assert len(frames) == 5
    for k, (c2w, fr) in enumerate(zip(poses, frames)):
        rays = r.generate_rays(c2w, H, W, focal, s=s, near=2.0, far=6.0)
        # (the rotation's multiply-adds round differently from the oracle's CPU matmul: same tolerance as the raygen golden test)
        assert torch.allclose(rays.cpu(), O.build_frame_rays(torch.from_numpy(c2w), H, W, focal, s, 2.0, 6.0), rtol=1e-5, atol=1e-5)
        out = r.forward_rays(rays, want_weights=False)
        with np.errstate(all="ignore"):
            ref_hr, ref_mat = F.assemble_frame(out["fine_comp_rgbs"].cpu().numpy(), out["fine_depth"].cpu().numpy(), H, W, s, 2.0, 6.0)
            lr_rgb = O.box_average(out["fine_comp_rgbs"].cpu(), s).numpy()
            lr_dep = O.box_average(out["fine_depth"].cpu(), s).numpy()
            ref_lr, _ = F.assemble_frame(lr_rgb, lr_dep, H // s, W // s, 1, 2.0, 6.0)
        assert np.array_equal(fr["fine_pred_ori"].numpy(), ref_hr), k
        assert np.array_equal(fr["fine_depth_mat_ori"].numpy(), ref_mat), k
        assert fr["fine_pred"].shape == (H // s, 2 * (W // s), 3)
        # the LR box average sums in a different order than torch.mean: allow one grey level on a few pixels
        diff = np.abs(fr["fine_pred"].numpy().astype(int) - ref_lr.astype(int))
        assert diff.max() <= 1 or (diff > 1).mean() < 1e-3, k
    assert not np.array_equal(frames[0]["fine_pred_ori"].numpy(), frames[2]["fine_pred_ori"].numpy())
    # forward-facing branch: NDC rays at near plane 1.0, near/far = 0/1 (data/llff_downX_dataset.py:473-481)
    sp = paths.spiral_poses(np.array([0.3, 0.2, 0.05]), 3.5, 4).astype(np.float32)
    res = r.render_test_pose(sp[1], H, W, focal, s, ndc=True)
    rays = r.generate_rays(sp[1], H, W, focal, s=s, ndc=True)
    assert torch.allclose(rays.cpu(), O.build_frame_rays(torch.from_numpy(sp[1]), H, W, focal, s, 0.0, 1.0, ndc=True), rtol=1e-5, atol=1e-5)
    out = r.forward_rays(rays, want_weights=False)
    with np.errstate(all="ignore"):
        ref_hr, _ = F.assemble_frame(out["coarse_comp_rgbs"].cpu().numpy(), out["coarse_depth"].cpu().numpy(), H, W, s, 0.0, 1.0)
    assert np.array_equal(res["coarse_pred_ori"].cpu().numpy(), ref_hr)
    r.close()
