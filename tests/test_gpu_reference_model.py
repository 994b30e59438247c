"""GPU: the reference's OWN model class under the one-line hook.

The unmodified `NeRFDownXModel` (models/nerf_downX_model.py) is built on cuda:0 from the staged copy of the
reference tree (tools/stage_reference.py -> baseline/_ref/NeRF-SR.tar.gz, unpacked by oracle/ref_shim.py; the
tests skip when no copy travelled to the box) and driven through its own methods, exactly as test.py / train.py do:

    inference   set_input -> forward() -> comp_low_res_output() -> calculate_vis()      (test.py:50-53, :316-353, :418-450)
    training    set_input -> optimize_parameters()                                      (train.py:70-80, :398-408)

first unpatched (stock PyTorch path, fp32, TF32 off), then after `patch_model(model)` -- same weights, same inputs,
same torch RNG seed -- and every `out_*` attribute, the losses, the gradients and the updated parameters are
compared.  Results go to gpurun_out/r02_reference_model.jsonl."""
import copy
import json
import math
import os

import pytest
import torch

from conftest import ROOT, e2e_bounds
from oracle import nerf_oracle as O
from oracle import ref_shim

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shim.reference_available(), reason="no copy of the reference tree on this box")]

COARSE = ("coarse_comp_rgbs", "coarse_depth", "coarse_opacity", "coarse_weights")
FINE = ("fine_comp_rgbs", "fine_depth", "fine_opacity", "fine_weights")


def _report(rec):
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "r02_reference_model.jsonl"), "a") as fh:
        fh.write(json.dumps(rec) + "\n")


def _psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 200.0 if mse == 0 else -10.0 * math.log10(mse)


def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _out_attrs(model):
    return {k[4:]: getattr(model, k).detach().clone() for k in sorted(vars(model)) if k.startswith("out_")
            and isinstance(getattr(model, k), torch.Tensor)}


def _inference(model, rays_grouped, rgbs, rgbs_ori):
    """test.py's per-frame body: model.set_input(data); model.test() == forward() under no_grad; then the
    validate() body (:473-479): calculate_losses() (-> comp_low_res_output) and calculate_vis(with_gt=True)."""
    model.eval()
    model.set_input({"rays": rays_grouped, "rgbs": rgbs, "rgbs_ori": rgbs_ori})
    with torch.no_grad():
        model.forward()
        model.calculate_losses()
        model.calculate_vis(with_gt=True)
    torch.cuda.synchronize()
    res = _out_attrs(model)
    for k in ("coarse_pred_img", "fine_pred_img", "coarse_pred_img_ori", "fine_pred_img_ori", "fine_depth_mat", "fine_depth_mat_ori"):
        res["vis_" + k] = getattr(model, k).detach().clone().cpu()
    res["loss_fine_psnr"] = torch.as_tensor(float(model.loss_fine_psnr))
    res["loss_coarse_psnr"] = torch.as_tensor(float(model.loss_coarse_psnr))
    return res


@pytest.mark.parametrize("kind,s,white,hw,seeds", [("blender", 2, True, (96, 128), (4, 17)), ("llff", 4, False, (64, 96), (21, 8))])
def test_real_reference_model_inference_unpatched_vs_patched(kind, s, white, hw, seeds):
    from nerf_sr_b200 import patch_model
    _no_tf32()
    H, W = hw
    n_lr = (H // s) * (W // s)
    args = ["--downscale", str(s), "--img_wh", str(W), str(H)] + (["--white_bkgd"] if white else [])
    model, opt = ref_shim.load_reference_model("nerf_downX", args, device="cuda:0")
    cfg = O.RenderConfig(white_bkgd=white, downscale=s)
    pc, pf = O.make_mlp_params(cfg, seeds[0]), O.make_mlp_params(cfg, seeds[1])
    ref_shim.set_weights(model, pc, pf)
    assert isinstance(model.netCoarse, torch.nn.DataParallel)        # init_net_dp with n_gpus=1 (models/networks.py:64-68)
    g = torch.Generator().manual_seed(1)
    rays = O.synthetic_rays(n_lr * s * s, 7, kind).view(1, n_lr, s * s, 8)     # the dataloader's [1, n_lr, s^2, 8] batch
    rgbs, rgbs_ori = torch.rand(1, n_lr, 3, generator=g), torch.rand(1, n_lr, s * s, 3, generator=g)

    ref = _inference(model, rays, rgbs, rgbs_ori)
    near_ref, far_ref = model.near.copy(), model.far.copy()
    assert patch_model(model, "bf16x3") is model and model._nsr_renderer is not None
    launches0 = model._nsr_renderer.launch_count
    got = _inference(model, rays, rgbs, rgbs_ori)
    assert model._nsr_renderer.launch_count > launches0              # the CUDA library rendered this frame
    assert type(model).__name__ == "NeRFDownXModel"
    assert model.near.shape == near_ref.shape and model.near.dtype == near_ref.dtype
    assert float(model.near[0]) == float(near_ref[0]) and float(model.far[0]) == float(far_ref[0])

    rec = {"test": "inference", "kind": kind, "s": s, "rays": n_lr * s * s, "keys": {}}
    assert set(got) == set(ref)
    for k in ref:
        assert got[k].shape == ref[k].shape and got[k].dtype == ref[k].dtype, k
        mx, v = O.tolerance_violations(got[k].float().cpu(), ref[k].float().cpu())
        rec["keys"][k] = {"max_abs": mx, "viol": v}
    # coarse stage (HR *_ori and box-averaged LR attributes): within tolerance everywhere -- except on a ray whose LAST coarse
    # sample has sigma within 1e-4 of zero: delta_last = 1e10 makes alpha_last a step function of sign(sigma_last)
    # (models/rendering.py:91-98), the one discontinuous decision of the coarse stage (tests/test_gpu_frame_parity.py)
    with torch.no_grad():
        ex = {}
        O.forward_rays({k: v.to("cuda:0") for k, v in pc.items()}, {k: v.to("cuda:0") for k, v in pf.items()}, rays.view(-1, 8).to("cuda:0"), cfg, extras=ex)
    well = (ex["raw_coarse"][:, -1, 3].abs() >= 1e-4).cpu()
    rec["ill_conditioned_rays_coarse"] = int((~well).sum())
    well_lr = well.view(n_lr, s * s).all(1)
    for k in ("coarse_comp_rgbs", "coarse_comp_rgbs_ori", "coarse_depth", "coarse_depth_ori", "coarse_opacity", "coarse_weights"):
        w = well if got[k].shape[0] == well.shape[0] else well_lr
        _, v = O.tolerance_violations(got[k].float().cpu()[w], ref[k].float().cpu()[w])
        assert v == 0.0, (k, rec["keys"][k], rec["ill_conditioned_rays_coarse"])
    assert rec["ill_conditioned_rays_coarse"] <= 3
    # fine stage end to end: ill-conditioned -> fp64 floor of the oracle on the same rays
    dev = torch.device("cuda:0")
    flat = rays.view(-1, 8).to(dev)
    with torch.no_grad():
        d = lambda p: {k: v.to(dev).double() for k, v in p.items()}
        ref64 = O.forward_rays(d(pc), d(pf), flat.double(), cfg)
    for k in FINE:
        src = k + "_ori" if k in ("fine_comp_rgbs", "fine_depth") else k          # HR tensors (the LR ones are box averages)
        _, floor = O.tolerance_violations(ref[src].cpu().reshape(ref64[k].shape), ref64[k].cpu())
        _, v64 = O.tolerance_violations(got[src].cpu().reshape(ref64[k].shape), ref64[k].cpu())
        rec["keys"][src]["floor_fp32_vs_fp64"] = floor
        rec["keys"][src]["viol_vs_fp64"] = v64
        b64, b32 = e2e_bounds(floor, n_lr * s * s, "bf16x3")
        assert v64 <= b64, (src, rec["keys"][src], b64)
        assert rec["keys"][src]["viol"] <= b32, (src, rec["keys"][src], b32)
    rec["psnr_fine_lr_image_db"] = _psnr(got["fine_comp_rgbs"], ref["fine_comp_rgbs"])
    rec["psnr_fine_vis_db"] = _psnr(got["vis_fine_pred_img"], ref["vis_fine_pred_img"])
    rec["loss_fine_psnr"] = [float(ref["loss_fine_psnr"]), float(got["loss_fine_psnr"])]
    _report(rec)
    assert rec["psnr_fine_lr_image_db"] > 50.0
    assert abs(float(got["loss_fine_psnr"]) - float(ref["loss_fine_psnr"])) < 1e-3      # the reported test metric
    assert got["vis_fine_pred_img"].shape == ref["vis_fine_pred_img"].shape            # [H/s, 3W/s, 3] pred | gt | depth
    model._nsr_renderer.close()


def test_real_reference_model_optimize_parameters_unpatched_vs_patched():
    """train.py's iteration (models/nerf_downX_model.py:398-408) on two models with identical weights and RNG seed:
    one stock, one patched.  Losses, all 48 gradient tensors and the Adam-updated parameters must agree."""
    from nerf_sr_b200 import patch_model
    _no_tf32()
    s, n_lr = 2, 256
    args = ["--downscale", "2", "--noise_std", "1.0", "--grad_clip_val", "0.1", "--use_var_loss"]
    cfg = O.RenderConfig(noise_std=1.0, downscale=s)
    pc, pf = O.make_mlp_params(cfg, 21), O.make_mlp_params(cfg, 8)
    rays = O.synthetic_rays(n_lr * s * s, 11, "llff").view(n_lr, s * s, 8)
    rgbs = torch.rand(n_lr, 3, generator=torch.Generator().manual_seed(3))

    def run(patched):
        model, opt = ref_shim.load_reference_model("nerf_downX", args, device="cuda:0", train=True)
        ref_shim.set_weights(model, pc, pf)
        if patched:
            patch_model(model, "bf16x3")
        losses, grads = [], None
        torch.manual_seed(1234)
        for it in range(2):
            model.set_input({"rays": rays.clone(), "rgbs": rgbs.clone()})
            model.optimize_parameters()
            losses.append({k: float(getattr(model, "loss_" + k)) for k in ("coarse_mse", "fine_mse", "tot", "out_coarse_var", "out_fine_var")})
            if it == 0:
                grads = [p.grad.detach().clone() for net in (model.netCoarse, model.netFine) for p in net.parameters()]
        params = [p.detach().clone() for net in (model.netCoarse, model.netFine) for p in net.parameters()]
        torch.cuda.synchronize()
        r = getattr(model, "_nsr_renderer", None)
        if patched:
            assert r is not None and r.launch_count > 0
            r.close()
        return losses, grads, params

    l_ref, g_ref, p_ref = run(False)
    l_got, g_got, p_got = run(True)
    rec = {"test": "optimize_parameters", "losses_ref": l_ref, "losses_patched": l_got, "grad_rel_l2": [], "grad_cos": []}
    for a, b in zip(g_got, g_ref):
        a, b = a.double().reshape(-1), b.double().reshape(-1)
        rec["grad_rel_l2"].append(float((a - b).norm() / (b.norm() + 1e-30)))
        rec["grad_cos"].append(float(torch.nn.functional.cosine_similarity(a[None], b[None])))
    lr = 5e-4                                                         # two Adam steps: each moves a parameter by ~lr at most
    rec["param_max_abs_diff_over_lr"] = max(float((a - b).abs().max()) for a, b in zip(p_got, p_ref)) / lr
    rec["param_mean_abs_diff_over_lr"] = max(float((a - b).abs().mean()) for a, b in zip(p_got, p_ref)) / lr
    _report(rec)
    # step 0 sees identical weights and identical draws: the coarse loss is a well-conditioned function of them
    assert abs(l_got[0]["coarse_mse"] - l_ref[0]["coarse_mse"]) <= 1e-4 * abs(l_ref[0]["coarse_mse"]) + 1e-7, rec
    for k in ("fine_mse", "tot", "out_coarse_var", "out_fine_var"):
        assert abs(l_got[0][k] - l_ref[0][k]) <= 2e-3 * abs(l_ref[0][k]) + 1e-6, (k, rec)
    for it in (1,):                                                   # after one update the trajectories still agree
        assert abs(l_got[it]["tot"] - l_ref[it]["tot"]) <= 1e-2 * abs(l_ref[it]["tot"]), rec
    assert min(rec["grad_cos"]) > 0.995, rec
    assert rec["param_max_abs_diff_over_lr"] <= 4.5 and rec["param_mean_abs_diff_over_lr"] < 0.5, rec
