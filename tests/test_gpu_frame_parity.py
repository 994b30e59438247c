"""GPU: full-frame parity at the BASELINE.json config sizes (SURVEY.md 8c iii).

C2 = configs[1]: 400x400 HR rays, 2x2 SS (160 000 rays, Blender-like); C4 = configs[3]: 800x800, 4x4 SS
(640 000 rays); C5 = configs[4]: 1008x756, 2x2 SS (762 048 rays, LLFF/NDC-like).  The checker is the oracle's
ATen op sequence run on cuda:0 in fp32 with TF32 off ("the bit-closest oracle for GPU sin/cos/exp"), in
the reference's own 4096-ray chunks.  Per frame:

  (i)   coarse stage, all rays:                    0 tolerance violations  (|a-b| <= 1e-4 + 1e-3 |b|)
  (ii)  fine stage teacher-forced on the oracle's fine z-values, all rays: 0 violations
  (iii) end to end (fine z-values recomputed from our own coarse weights): violation fraction and PSNR of the
        LR image (after the s x s box average) and of the HR image against the oracle's.  Fine sample positions are an
        ill-conditioned function of the coarse weights, so the pass criterion is relative to the
        oracle's own fp32-vs-fp64 disagreement ("floor") measured on the first FLOOR_RAYS rays of the same frame:
            viol(ours vs fp64)  <= viol(oracle fp32 vs fp64) + 0.01      -- we are as close to the exact answer as the
                                                                            reference's own arithmetic is
            viol(ours vs oracle fp32) <= 2 * floor + 0.01                -- two independent fp32-grade evaluations
        and the LR-image PSNR against the oracle must exceed 50 dB.

Every number is appended to gpurun_out/r02_frame_parity.jsonl (copied to profiles/r02_frame_parity.md)."""
import json
import math
import os

import pytest
import torch

from conftest import ROOT
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu

FLOOR_RAYS = 32768
FRAMES = {
    # name: (rays, s, kind, white_bkgd, seeds)
    "C2_blender_400x400_s2": (160000, 2, "blender", True, (4, 17)),
    "C4_blender_800x800_s4": (640000, 4, "blender", True, (4, 17)),
    "C5_llff_1008x756_s2": (762048, 2, "llff", False, (21, 8)),
}
COARSE = ("coarse_comp_rgbs", "coarse_depth", "coarse_opacity", "coarse_weights")
FINE = ("fine_comp_rgbs", "fine_depth", "fine_opacity", "fine_weights")


def _report(rec):
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "r02_frame_parity.jsonl"), "a") as fh:
        fh.write(json.dumps(rec) + "\n")


def _psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 200.0 if mse == 0 else -10.0 * math.log10(mse)


def _oracle_frame(pc, pf, rays, cfg, chunk=4096, dtype=torch.float32):
    """chunk_batch(forward_rays, 4096, rays) (models/nerf_downX_model.py:318) with the fine z-values kept."""
    cast = lambda t: t.to(dtype)
    pc, pf = {k: cast(v) for k, v in pc.items()}, {k: cast(v) for k, v in pf.items()}
    acc = {}
    with torch.no_grad():
        for i in range(0, rays.shape[0], chunk):
            ex = {}
            out = O.forward_rays(pc, pf, cast(rays[i:i + chunk]), cfg, extras=ex)
            out["z_fine"] = ex["z_fine"]
            for k, v in out.items():
                acc.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in acc.items()}


def _viol(a, b):
    mx, v = O.tolerance_violations(a, b)
    return mx, v


@pytest.mark.parametrize("prec", ["bf16x3"])
@pytest.mark.parametrize("name", list(FRAMES))
def test_full_frame_against_the_oracle_on_the_gpu(name, prec):
    from nerf_sr_b200 import Renderer
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    n, s, kind, white, seeds = FRAMES[name]
    dev = torch.device("cuda:0")
    cfg = O.RenderConfig(white_bkgd=white, downscale=s)
    pc, pf = O.make_mlp_params(cfg, seeds[0]), O.make_mlp_params(cfg, seeds[1])
    rays = O.synthetic_rays(n, 900 + s, kind).to(dev)
    pcd, pfd = {k: v.to(dev) for k, v in pc.items()}, {k: v.to(dev) for k, v in pf.items()}
    ref = _oracle_frame(pcd, pfd, rays, cfg)
    assert 0.05 < float(ref["coarse_opacity"].mean()) < 0.999 and 0.05 < float(ref["fine_opacity"].mean()) < 0.999

    r = Renderer(cfg, dev, precision=prec)
    r.load_state_dict(0, pc)
    r.load_state_dict(1, pf)
    out = r.forward_rays(rays)
    rec = {"frame": name, "precision": prec, "rays": n, "s": s, "stages": {}}
    # (i) coarse stage
    for k in COARSE:
        mx, v = _viol(out[k], ref[k])
        rec["stages"][k] = {"max_abs": mx, "viol": v}
        assert v == 0.0, (k, mx, v)
    # (ii) fine stage, teacher-forced on the oracle's z-values
    tf = r.render_pass(1, rays, ref["z_fine"])
    for kl, kr in (("comp_rgbs", "fine_comp_rgbs"), ("depth", "fine_depth"), ("opacity", "fine_opacity"), ("weights", "fine_weights")):
        mx, v = _viol(tf[kl], ref[kr])
        rec["stages"]["teacher_forced_" + kr] = {"max_abs": mx, "viol": v}
        assert v == 0.0, (kr, mx, v)
    del tf
    # (iii) end to end
    m = min(FLOOR_RAYS, n)
    ref64 = _oracle_frame(pcd, pfd, rays[:m], cfg, chunk=2048, dtype=torch.float64)
    for k in FINE:
        mx, v = _viol(out[k], ref[k])
        _, floor = _viol(ref[k][:m], ref64[k])
        mx64, v64 = _viol(out[k][:m], ref64[k])
        _, v32 = _viol(out[k][:m], ref[k][:m])
        rec["stages"]["e2e_" + k] = {"max_abs": mx, "viol": v, "floor_fp32_vs_fp64": floor, "viol_vs_fp64": v64,
                                     "max_abs_vs_fp64": mx64, "viol_vs_fp32_on_floor_rays": v32}
    lr_ours, lr_ref = r.box_average(out["fine_comp_rgbs"], s), O.box_average(ref["fine_comp_rgbs"], s)
    rec["psnr_lr_fine_vs_oracle_db"] = _psnr(lr_ours, lr_ref)
    rec["psnr_hr_fine_vs_oracle_db"] = _psnr(out["fine_comp_rgbs"], ref["fine_comp_rgbs"])
    rec["psnr_lr_coarse_vs_oracle_db"] = _psnr(r.box_average(out["coarse_comp_rgbs"], s), O.box_average(ref["coarse_comp_rgbs"], s))
    rec["psnr_lr_oracle_fp32_vs_fp64_db"] = _psnr(O.box_average(ref["fine_comp_rgbs"][:m], s), O.box_average(ref64["fine_comp_rgbs"], s))
    rec["psnr_lr_ours_vs_fp64_db"] = _psnr(lr_ours[: m // (s * s)], O.box_average(ref64["fine_comp_rgbs"], s))
    _report(rec)
    for k in FINE:
        st = rec["stages"]["e2e_" + k]
        assert st["viol_vs_fp64"] <= st["floor_fp32_vs_fp64"] + 0.01, (k, st)
        assert st["viol"] <= 2.0 * st["floor_fp32_vs_fp64"] + 0.01, (k, st)
        assert torch.isfinite(out[k]).all()
    assert rec["psnr_lr_fine_vs_oracle_db"] > 50.0, rec
    r.close()
