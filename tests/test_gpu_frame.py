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
