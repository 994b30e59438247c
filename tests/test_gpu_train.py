"""GPU tests of the training path (scope row f-1), all through the C ABI: activation stash of the fused
forward, the tcgen05 backward GEMMs (dX, dW) against torch fp64 matmuls, full gradients against the
oracle's autograd (= the reference's definition, pinned in tests/golden/train_step_*.npz), the LR loss +
its gradient, clipping, Adam against torch.optim.Adam, and the fused training iteration.

Gradient parity: the fine sample positions are an ill-conditioned function of the coarse weights
(SURVEY.md 0.6), so the oracle is teacher-forced on the z-values the CUDA forward actually used
(protocol ii); what remains is compared as relative L2 per parameter tensor against the oracle's own
fp32-vs-fp64 disagreement on the same inputs (ReLU / sigma-ReLU kink flips make that floor ~1e-3..1e-2)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import TrainFixture, rng_dict, train_golden_names
from oracle import nerf_oracle as O
from oracle import train_oracle as T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "train_parity.jsonl")


def _report(**kw):
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as fh:
            fh.write(json.dumps(kw) + "\n")
    except OSError:
        pass


def _renderer(cfg, pc, pf, prec="bf16x3"):
    from nerf_sr_b200 import Renderer
    r = Renderer(cfg, torch.device(DEV), precision=prec)
    r.load_state_dict(0, pc)
    r.load_state_dict(1, pf)
    return r


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(torch.linalg.vector_norm(a - b) / (torch.linalg.vector_norm(b) + 1e-30))


def _split_eps(prec):
    return 2.0 ** -15 if prec == "bf16x3" else 2.0 ** -20


# ----------------------------------------------------------------------------- tile images
@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3"])
def test_image_pack_roundtrip(prec):
    cfg = O.RenderConfig()
    r = _renderer(cfg, O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17), prec)
    g = torch.Generator().manual_seed(0)
    for rows, ld in ((300, 256), (128, 63), (1, 128), (1000, 27)):
        x = torch.randn(rows, ld, generator=g).to(DEV)
        img = r.pack_image(x)
        n_cols = (ld + 63) // 64 * 64
        y = r.unpack_image(img, rows, n_cols, ld)
        assert float((x - y).abs().max()) <= _split_eps(prec) * float(x.abs().max()), (rows, ld)
    r.close()


# ----------------------------------------------------------------------------- activation stash
@pytest.mark.parametrize("prec", ["bf16x3"])
def test_stash_matches_oracle_activations(prec):
    cfg = O.RenderConfig(white_bkgd=True)
    pc, pf = O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17)
    n = 101                                         # ragged: last coarse tile half empty
    rays = O.synthetic_rays(n, 3, "blender")
    r = _renderer(cfg, pc, pf, prec)
    out = r.render_train(rays.to(DEV), None, want_z_fine=True)
    ref_eval = r.forward_rays(rays.to(DEV))          # the non-stash kernel must agree bit for bit
    torch.cuda.synchronize()
    for k in ("coarse_comp_rgbs", "coarse_weights", "fine_comp_rgbs", "fine_depth"):
        assert torch.equal(out[k], ref_eval[k]), k
    z_f = out["z_fine"].cpu()
    extras = {}
    with torch.no_grad():
        O.forward_rays(pc, pf, rays, cfg, None, z_fine_override=z_f, extras=extras)
    for which, (p, z) in enumerate(((pc, extras["z_coarse"]), (pf, z_f))):
        xyz = O.cast_rays(rays[:, 0:3], rays[:, 3:6], z).reshape(-1, 3)
        enc = O.posenc(xyz, cfg.deg_pos)
        denc = O.posenc(rays[:, 3:6], cfg.deg_dir).repeat_interleave(z.shape[1], dim=0)
        acts = []
        with torch.no_grad():
            O.mlp_forward(p, torch.cat([enc, denc], -1), cfg, acts=acts)
        got = r.stash_activation(n, which, 0).cpu()
        mx, viol = O.tolerance_violations(got[:, :63], enc)
        assert viol == 0.0 and float(got[:, 63].abs().max()) == 0.0, ("enc", which, mx)
        for layer, ref in enumerate(acts, start=1):       # h1..h8, feat, dir
            got = r.stash_activation(n, which, layer).cpu()
            mx, viol = O.tolerance_violations(got, ref)
            _report(test="stash", prec=prec, which=which, layer=layer, max_abs=mx, viol=viol)
            assert viol == 0.0, (which, layer, mx, viol)
            if layer <= 8:      # the 1-bit ReLU mask the dX kernels use == (stashed activation > 0)
                assert torch.equal(r.stash_mask(n, which, layer).cpu(), got > 0), (which, layer)
    r.close()


# ----------------------------------------------------------------------------- dX GEMM
@pytest.mark.parametrize("prec", ["bf16x3"])
def test_dx_gemm_against_torch(prec):
    cfg = O.RenderConfig()
    pc, pf = O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17)
    r = _renderer(cfg, pc, pf, prec)
    g = torch.Generator().manual_seed(1)
    names = ["dir_encoding.0.weight", "xyz_encoding_final.weight"] + [f"xyz_encoding_{L}.0.weight" for L in range(8, 1, -1)]
    for rows in (300, 148 * 128 + 77):
        for idx, name in enumerate(names):
            if rows > 1000 and idx not in (0, 1, 5):
                continue
            W = pf[name].double()
            k_out = W.shape[0]
            col0 = 63 if name == "xyz_encoding_5.0.weight" else 0
            dz = torch.randn(rows, k_out, generator=g)
            hmask = torch.relu(torch.randn(rows, 256, generator=g))
            a_img = r.pack_image(dz.to(DEV))
            m_bits = r.relu_bits(hmask.to(DEV))
            dsig = wsig = None
            ref = dz.double() @ W[:, col0:col0 + 256]
            if idx == 1:
                dsig = torch.randn((rows + 127) // 128 * 128, generator=g)
                dsig[rows:] = 0
                wsig = torch.randn(256, generator=g)
                ref = ref + dsig[:rows, None].double() * wsig[None].double()
                dsig, wsig = dsig.to(DEV), wsig.to(DEV)
            use_mask = idx != 0
            if use_mask:
                ref = ref * (hmask > 0)
            out_img = r.debug_dx(1, idx, a_img, rows, m_bits if use_mask else None, dsig, wsig)
            got = r.unpack_image(out_img, rows, 256).cpu().double()
            err = float((got - ref).abs().max()) / float(ref.abs().max())
            _report(test="dx", prec=prec, rows=rows, layer=name, rel_max_err=err)
            assert err < (1e-4 if prec == "bf16x3" else 2e-5), (name, rows, err)
    r.close()


# ----------------------------------------------------------------------------- dW GEMM
@pytest.mark.parametrize("prec", ["bf16x3"])
def test_dw_gemm_against_torch(prec):
    cfg = O.RenderConfig()
    r = _renderer(cfg, O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17), prec)
    g = torch.Generator().manual_seed(2)
    for rows, a_cols, blk, b_cols in ((300, 256, (0, 1), 256), (300, 256, (2, 3), 64), (1000, 128, (0, 1), 128),
                                      (148 * 128 * 2 + 500, 256, (2, 3), 256), (5000, 64, (0, 0), 256), (40, 256, (1, 2), 192)):
        A = torch.randn(rows, a_cols, generator=g)
        B = torch.randn(rows, b_cols, generator=g)
        out, bias = r.debug_dw(r.pack_image(A.to(DEV)), a_cols, blk[0], blk[1], r.pack_image(B.to(DEV)), b_cols, rows)
        sel = torch.cat([A[:, 64 * blk[0]:64 * blk[0] + 64], A[:, 64 * blk[1]:64 * blk[1] + 64]], 1).double()
        ref = sel.T @ B.double()
        err = float((out.cpu().double() - ref).abs().max()) / float(ref.abs().max())
        berr = float((bias.cpu().double() - sel.sum(0)).abs().max()) / float(sel.sum(0).abs().max())
        _report(test="dw", prec=prec, rows=rows, a_cols=a_cols, blk=blk, b_cols=b_cols, rel_max_err=err, bias_err=berr)
        assert err < 1e-4 and berr < 1e-4, (rows, a_cols, blk, b_cols, err, berr)
    r.close()


# ----------------------------------------------------------------------------- loss, clip, Adam
@pytest.mark.parametrize("s,lam", [(2, 1.0), (4, 0.5), (1, 1.0)])
def test_lr_loss_grad_against_autograd(s, lam):
    cfg = O.RenderConfig()
    r = _renderer(cfg, O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17))
    g = torch.Generator().manual_seed(5)
    n_lr = 777
    hr = torch.rand(n_lr * s * s, 3, generator=g, requires_grad=True)
    tgt = torch.rand(n_lr, 3, generator=g)
    lr_ref = O.box_average(hr, s)
    loss = torch.nn.functional.mse_loss(lr_ref, tgt) * lam
    loss.backward()
    psnr = -10 * torch.log10(torch.mean((lr_ref.detach() - tgt) ** 2))
    lr, m, ghr = r.lr_loss_grad(hr.detach().to(DEV), tgt.to(DEV), s, lam)
    assert torch.allclose(lr.cpu(), lr_ref.detach(), rtol=0, atol=5e-7)      # torch.mean sums in a different order
    assert float(m[0]) == pytest.approx(float(loss.detach()), rel=1e-5) and float(m[1]) == pytest.approx(float(psnr), rel=1e-5)
    scale = 2.0 * lam / (3 * n_lr) / (s * s)      # lr differs from torch.mean by ~1e-7 where the sum order differs
    assert torch.allclose(ghr.cpu(), hr.grad, rtol=1e-5, atol=1e-6 * scale)
    r.close()


def test_clip_and_adam_against_torch():
    cfg = O.RenderConfig()
    pc = O.make_mlp_params(cfg, 4)
    r = _renderer(cfg, pc, O.make_mlp_params(cfg, 17))
    names = list(pc)
    g = torch.Generator().manual_seed(6)
    ref = [torch.nn.Parameter(pc[n].clone()) for n in names]
    opt = torch.optim.Adam(ref, lr=1e-3, betas=(0.9, 0.999))
    mine = [pc[n].clone().to(DEV) for n in names]
    numel = sum(p.numel() for p in mine)
    m, v = torch.zeros(numel, device=DEV), torch.zeros(numel, device=DEV)
    for step in range(1, 4):
        grads = [torch.randn(p.shape, generator=g) * 10 ** float(2 * torch.randn((), generator=g)) for p in ref]
        flat = torch.cat([x.reshape(-1) for x in grads]).to(DEV)
        for p, x in zip(ref, grads):
            p.grad = x.clone()
        total = torch.nn.utils.clip_grad_norm_(ref, 0.5)
        coef = r.clip_coef(flat, None, 0.5)
        assert float(coef[1]) == pytest.approx(float(total), rel=1e-5)
        opt.step()
        r.adam_step(mine, flat, m, v, step, 1e-3, clip_coef_dev=coef)
        torch.cuda.synchronize()
        for a, b, n in zip(mine, ref, names):
            assert torch.allclose(a.cpu(), b.detach(), rtol=2e-6, atol=2e-9), (step, n, float((a.cpu() - b.detach()).abs().max()))
    # value clipping
    ref2 = [torch.nn.Parameter(pc[n].clone()) for n in names]
    opt2 = torch.optim.Adam(ref2, lr=5e-4)
    mine2 = [pc[n].clone().to(DEV) for n in names]
    m.zero_(); v.zero_()
    grads = [torch.randn(p.shape, generator=g) for p in ref2]
    for p, x in zip(ref2, grads):
        p.grad = x.clone()
    torch.nn.utils.clip_grad_value_(ref2, 0.2)
    opt2.step()
    r.adam_step(mine2, torch.cat([x.reshape(-1) for x in grads]).to(DEV), m, v, 1, 5e-4, clip_value=0.2)
    for a, b in zip(mine2, ref2):
        assert torch.allclose(a.cpu(), b.detach(), rtol=2e-6, atol=2e-9)
    r.close()


@pytest.mark.parametrize("s,n_lr,terms", [(2, 777, "all"), (4, 1001, "all"), (3, 50, "all"), (2, 70001, "var"), (2, 300, "sr"),
                                          (2, 300, "mse"), (1, 64, "sr"), (2, 300, "ref"), (4, 33, "ref")])
def test_loss_epilogue_against_autograd(s, n_lr, terms):
    """nsr_loss_epilogue: every term of calculate_losses (box average + lambda*MSE + PSNR, sub-pixel variance of colour
    and of depth/far, SISR MSE) and the gradient to the HR outputs, against torch autograd over the reference's
    expressions (models/nerf_downX_model.py:326-378; restated in oracle.train_oracle.subpixel_variance_sum)."""
    cfg = O.RenderConfig()
    r = _renderer(cfg, O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17))
    g = torch.Generator().manual_seed(5 + s)
    n = n_lr * s * s
    hr = torch.rand(n, 3, generator=g, dtype=torch.float64, requires_grad=True)
    depth = (2 + 4 * torch.rand(n, generator=g, dtype=torch.float64)).requires_grad_(True)
    tgt = torch.rand(n_lr, 3, generator=g, dtype=torch.float64)
    tgt_hr = torch.rand(n, 3, generator=g, dtype=torch.float64) if terms in ("all", "sr", "ref") else None
    lam, lam_v, lam_d, far = 0.7, (0.02 if terms in ("all", "var") else 0.0), (0.05 if terms in ("all", "var") else 0.0), 6.0
    lam_hr = 1.0 / (s * s) if terms == "ref" else 1.0        # --with_ref batch: no LR target, HR MSE / s^2
    lr_ref = O.box_average(hr, s)
    if terms == "ref":
        tot, ref = 0.0, {"mse": 0.0, "psnr": 0.0, "var": 0.0, "dvar": 0.0, "sr": 0.0}
    else:
        mse = torch.nn.functional.mse_loss(lr_ref, tgt)
        tot = mse * lam
        ref = {"mse": float(mse.detach() * lam), "psnr": float(-10 * torch.log10(mse.detach())), "var": 0.0, "dvar": 0.0, "sr": 0.0}
    if tgt_hr is not None:
        sr = torch.nn.functional.mse_loss(hr, tgt_hr) * lam_hr
        tot = tot + sr
        ref["sr"] = float(sr.detach())
    if lam_v:
        v = T.subpixel_variance_sum(hr, n_lr, s)
        tot = tot + lam_v * v
        ref["var"] = float(v.detach())
    if lam_d:
        dv = T.subpixel_variance_sum(depth, n_lr, s, far)
        tot = tot + lam_d * dv
        ref["dvar"] = float(dv.detach())
    tot.backward()
    f = lambda t: None if t is None else t.detach().float().to(DEV)
    e = r.loss_epilogue(f(hr), None if terms == "ref" else f(tgt), s, lam, hr_depth=f(depth), lambda_var=lam_v, lambda_depth_var=lam_d,
                        far=far, target_hr=f(tgt_hr), lambda_hr=lam_hr)
    torch.cuda.synchronize()
    m = e["metrics"].cpu()
    assert torch.allclose(e["lr_rgb"].cpu().double(), lr_ref.detach(), rtol=0, atol=5e-7)
    assert torch.allclose(e["lr_depth"].cpu().double(), O.box_average(depth.detach(), s).reshape(-1), rtol=0, atol=2e-6)
    for i, k in enumerate(("mse", "psnr", "var", "dvar", "sr")):
        assert float(m[i]) == pytest.approx(ref[k], rel=2e-5, abs=1e-9), (k, float(m[i]), ref[k])
    assert float(m[5]) == pytest.approx(float(tot.detach()), rel=2e-5)
    g_rgb, g_depth = e["g_rgb"].cpu().double(), e["g_depth"].cpu().double()
    assert torch.allclose(g_rgb, hr.grad, rtol=1e-4, atol=1e-6 * float(hr.grad.abs().max())), float((g_rgb - hr.grad).abs().max())
    gd_ref = depth.grad if depth.grad is not None else torch.zeros(n, dtype=torch.float64)
    assert torch.allclose(g_depth, gd_ref, rtol=1e-4, atol=5e-6 * float(gd_ref.abs().max()) + 1e-30), float((g_depth - gd_ref).abs().max())
    # errors: variance terms at s = 1, depth term without depth
    from nerf_sr_b200 import NsrError
    if s == 1:
        with pytest.raises(NsrError):
            r.loss_epilogue(f(hr), f(tgt), 1, 1.0, lambda_var=0.01)
    with pytest.raises(NsrError):
        r.loss_epilogue(f(hr), f(tgt), s, 1.0, lambda_depth_var=0.01, far=6.0)
    with pytest.raises(NsrError):
        r.loss_epilogue(f(hr), None, s, 1.0)                      # no target at all
    r.close()


# ----------------------------------------------------------------------------- full gradients
def _trainer_kwargs(fx):
    """Trainer arguments for a fixture's reference flags (lambda_*_var count only when --use_*_loss is given)."""
    t = fx.tcfg
    return dict(lambda_coarse_mse=t.lambda_coarse_mse, lambda_fine_mse=t.lambda_fine_mse, downscale=fx.s,
                lambda_coarse_var=t.lambda_coarse_var if t.use_var_loss else 0.0,
                lambda_fine_var=t.lambda_fine_var if t.use_var_loss else 0.0,
                lambda_coarse_depth_var=t.lambda_coarse_depth_var if t.use_depth_var_loss else 0.0,
                lambda_fine_depth_var=t.lambda_fine_depth_var if t.use_depth_var_loss else 0.0)


def _oracle_grads(fx, z_f, rng, dtype=torch.float32, ref_z_f=None):
    cast = lambda t: None if t is None else t.to(dtype)
    conv = lambda g: None if g is None else O.RenderRng(cast(g.u_coarse), cast(g.noise_coarse), cast(g.u_fine), cast(g.noise_fine))
    rr = conv(rng)
    pc = {k: v.to(dtype) for k, v in fx.p_coarse.items()}
    pf = {k: v.to(dtype) for k, v in fx.p_fine.items()}
    return T.loss_and_grads(pc, pf, fx.rays.to(dtype), fx.target.to(dtype), fx.cfg, fx.tcfg, rr, fx.s, z_fine_override=z_f.to(dtype),
                            target_sr=None if fx.target_sr is None else fx.target_sr.to(dtype),
                            ref_rays=cast(fx.ref_rays), ref_rgbs=cast(fx.ref_rgbs), ref_rng=conv(fx.ref_rng[0]),
                            ref_z_fine_override=cast(ref_z_f))


@pytest.mark.parametrize("prec", ["bf16x3"])
@pytest.mark.parametrize("name", train_golden_names())
def test_gradients_against_oracle_autograd(name, prec):
    from nerf_sr_b200 import Trainer
    fx = TrainFixture(name)
    r = _renderer(fx.cfg, fx.p_coarse, fx.p_fine, prec)
    tr = Trainer(r, fx.p_coarse, fx.p_fine, **_trainer_kwargs(fx))
    rng = rng_dict(fx.rng[0])
    rays = fx.rays.to(DEV)
    out = r.render_train(rays, rng, want_z_fine=True)
    z_f = out["z_fine"].cpu()
    dev = lambda t: None if t is None else t.to(DEV)
    ref_rng, ref_z_f = rng_dict(fx.ref_rng[0]), None
    if fx.ref_rays is not None:             # --with_ref: the second forward is teacher-forced on its own CUDA z-values too
        ref_z_f = r.render_train(dev(fx.ref_rays), ref_rng, want_z_fine=True)["z_fine"].cpu()
    gc, gf = tr.forward_backward(rays, fx.target.to(DEV), rng, target_sr=dev(fx.target_sr), far=float(fx.rays[0, 7]),
                                 ref_rays=dev(fx.ref_rays), ref_rgbs=dev(fx.ref_rgbs), ref_rng=ref_rng)
    torch.cuda.synchronize()
    assert torch.isfinite(gc).all() and torch.isfinite(gf).all()
    losses, oc, of, _ = _oracle_grads(fx, z_f, fx.rng[0], ref_z_f=ref_z_f)
    _, oc64, of64, _ = _oracle_grads(fx, z_f, fx.rng[0], torch.float64, ref_z_f=ref_z_f)
    m = tr.last_metrics.cpu()
    assert float(m[0]) == pytest.approx(float(losses["coarse_mse"]), rel=2e-3)
    assert float(m[2]) == pytest.approx(float(losses["fine_mse"]), rel=2e-3)
    assert float(m[1]) == pytest.approx(float(losses["coarse_psnr"]), rel=2e-3)
    if tr.last_terms is not None:           # the sub-pixel variance / SISR terms of the fused epilogue
        lt = tr.last_terms.cpu()
        for w, net in enumerate(("coarse", "fine")):
            for col, key in ((2, f"{net}_var"), (3, f"{net}_depth_var"), (4, f"{net}_mse_sr")):
                if key in losses:
                    _report(test="loss_terms", fixture=name, term=key, got=float(lt[w, col]), oracle=float(losses[key]))
                    assert float(lt[w, col]) == pytest.approx(float(losses[key]), rel=5e-3), (key, float(lt[w, col]), float(losses[key]))
        tot = float(lt[0, 5] + lt[1, 5])
        if fx.ref_rays is not None:
            rt = tr.last_ref_terms.cpu()
            assert float(rt[0]) == pytest.approx(float(losses["ref_coarse_mse"]), rel=5e-3)
            assert float(rt[1]) == pytest.approx(float(losses["ref_fine_mse"]), rel=5e-3)
            tot += float(rt.sum())
        assert tot == pytest.approx(float(losses["tot"]), rel=3e-3)
    worst = 0.0
    for net, flat, ref32, ref64 in (("coarse", gc, oc, oc64), ("fine", gf, of, of64)):
        off = 0
        for k, ref in ref32.items():
            n = ref.numel()
            got = flat[off:off + n].view_as(ref).cpu()
            off += n
            floor = _rel(ref, ref64[k])
            err = _rel(got, ref64[k])
            cos = float(torch.nn.functional.cosine_similarity(got.double().reshape(1, -1), ref64[k].reshape(1, -1)))
            _report(test="grads", fixture=name, prec=prec, net=net, param=k, rel_l2=err, fp32_floor=floor, cos=cos,
                    ref_norm=float(torch.linalg.vector_norm(ref64[k])))
            worst = max(worst, err / max(floor, 1e-4))
            # bf16x3 pre-activations differ from fp32 by ~1e-5, so ~10x more ReLU masks flip than between fp32
            # and fp64; measured (profiles/r01_train_parity.md): <= 2.5e-2 on the first layer, <= 1e-3 from L6 up
            assert err <= 3.0 * floor + 3e-2, (net, k, err, floor)
            assert cos > 0.9995, (net, k, cos)
            if k.startswith(('rgb', 'sigma', 'dir_encoding', 'xyz_encoding_final')):
                assert err <= 3.0 * floor + 5e-3, (net, k, err, floor)    # a single relu(sigma) flip shows at ~2e-3 on 192 rays
        assert off == flat.numel()
    r.close()


@pytest.mark.parametrize("name", ["train_step_blender", "train_step_llff_clip", "train_step_s4_value_clip"])
def test_gradients_with_relu_masks_teacher_forced(name):
    """VERDICT r1 weak #1 / ADVICE: isolate the GEMM + compositing backward from ReLU decisions.  The free-running test
    above attributes its 1e-2-level first-layer residual to ReLU masks that flip on ~1e-5 pre-activation differences;
    here the oracle's autograd is run with the masks the CUDA forward actually stashed (8 trunk layers + the dir layer of
    both nets) and with the CUDA forward's relu(sigma) decisions, so every gradient must sit at the arithmetic floor:
    rel-L2 <= 3 x (fp32-vs-fp64 oracle) + 2e-4 (measured: 2e-6 .. 3e-5)."""
    from nerf_sr_b200 import Trainer
    fx = TrainFixture(name)
    r = _renderer(fx.cfg, fx.p_coarse, fx.p_fine, "bf16x3")
    tr = Trainer(r, fx.p_coarse, fx.p_fine, **_trainer_kwargs(fx))
    rng = rng_dict(fx.rng[0])
    rays = fx.rays.to(DEV)
    n = rays.shape[0]
    z_f = r.render_train(rays, rng, want_z_fine=True)["z_fine"].cpu()
    gc, gf = tr.forward_backward(rays, fx.target.to(DEV), rng, far=float(fx.rays[0, 7]))
    torch.cuda.synchronize()
    masks = [[r.stash_mask(n, w, l).cpu() for l in range(1, 9)] + [(r.stash_activation(n, w, 10) > 0).cpu()] for w in (0, 1)]
    # ... and the sign decisions of relu(sigma) in the compositing (the last sample's delta is 1e10: alpha jumps 0 <-> 1):
    # sigma as the CUDA forward computed it (the non-stash kernel is bit-identical, test_stash_matches_oracle_activations)
    z_c = r.sample_along_rays(rays, rng["u_coarse"].to(DEV) if (rng and rng.get("u_coarse") is not None) else None)
    for w, z, nk in ((0, z_c, "noise_coarse"), (1, z_f.to(DEV), "noise_fine")):
        nz = rng.get(nk).to(DEV) if (rng and rng.get(nk) is not None and fx.cfg.noise_std > 0) else None
        sig = r.render_pass(w, rays, z, nz, want_raw=True)["raw"][..., 3]
        if nz is not None:
            sig = sig + nz * fx.cfg.noise_std
        masks[w].append((sig > 0).cpu())
    masks = tuple(masks)

    def oracle(dtype):
        cast = lambda t: None if t is None else t.to(dtype)
        g = fx.rng[0]
        rr = O.RenderRng(cast(g.u_coarse), cast(g.noise_coarse), cast(g.u_fine), cast(g.noise_fine))
        return T.loss_and_grads({k: v.to(dtype) for k, v in fx.p_coarse.items()}, {k: v.to(dtype) for k, v in fx.p_fine.items()},
                                fx.rays.to(dtype), fx.target.to(dtype), fx.cfg, fx.tcfg, rr, fx.s, z_fine_override=z_f.to(dtype),
                                relu_masks=masks)
    _, oc, of, _ = oracle(torch.float32)
    _, oc64, of64, _ = oracle(torch.float64)
    worst = 0.0
    for net, flat, ref32, ref64 in (("coarse", gc, oc, oc64), ("fine", gf, of, of64)):
        off = 0
        for k, ref in ref32.items():
            m = ref.numel()
            got = flat[off:off + m].view_as(ref).cpu()
            off += m
            floor, err = _rel(ref, ref64[k]), _rel(got, ref64[k])
            _report(test="grads_mask_teacher_forced", fixture=name, net=net, param=k, rel_l2=err, fp32_floor=floor)
            worst = max(worst, err)
            assert err <= 3.0 * floor + 1e-3, (net, k, err, floor)
    r.close()


def test_autograd_function_matches_trainer_and_reference_adam():
    """RenderFunction: loss.backward() fills p.grad with the same gradients Trainer computes, and torch's
    own Adam on those parameters matches nsr_adam_step (the drop-in path of patch_model in train mode)."""
    from nerf_sr_b200 import RenderFunction, Trainer
    from nerf_sr_b200.training import OUT_KEYS
    fx = TrainFixture("train_step_blender")
    r = _renderer(fx.cfg, fx.p_coarse, fx.p_fine)
    names = list(fx.p_coarse)
    pcs = [torch.nn.Parameter(fx.p_coarse[n].clone().to(DEV)) for n in names]
    pfs = [torch.nn.Parameter(fx.p_fine[n].clone().to(DEV)) for n in names]
    rng = rng_dict(fx.rng[0])
    rays, tgt = fx.rays.to(DEV), fx.target.to(DEV)
    outs = dict(zip(OUT_KEYS, RenderFunction.apply(r, rays, rng, len(pcs), None, *pcs, *pfs)))
    lr_c = O.box_average(outs["coarse_comp_rgbs"], fx.s)
    lr_f = O.box_average(outs["fine_comp_rgbs"], fx.s)
    loss = torch.nn.functional.mse_loss(lr_c, tgt) + torch.nn.functional.mse_loss(lr_f, tgt)
    loss.backward()
    tr = Trainer(r, fx.p_coarse, fx.p_fine, downscale=fx.s)
    gc, gf = tr.forward_backward(rays, tgt, rng)
    flat_c = torch.cat([p.grad.reshape(-1) for p in pcs])
    flat_f = torch.cat([p.grad.reshape(-1) for p in pfs])
    assert _rel(flat_c, gc) < 1e-5 and _rel(flat_f, gf) < 1e-5
    assert float(loss) == pytest.approx(float(tr.last_metrics[0] + tr.last_metrics[2]), rel=1e-5)
    opt = torch.optim.Adam(pcs + pfs, lr=5e-4, betas=(0.9, 0.999))
    opt.step()
    tr2 = Trainer(r, fx.p_coarse, fx.p_fine, downscale=fx.s)
    tr2.optimize_parameters(rays, tgt, rng)
    for a, b in zip(pcs + pfs, tr2.params[0] + tr2.params[1]):
        assert torch.allclose(a.detach(), b, rtol=2e-6, atol=1e-8), float((a.detach() - b).abs().max())
    r.close()


def test_depth_and_opacity_gradients():
    """dL/d(depth) and dL/d(opacity) (used by the reference's depth-variance term) against autograd."""
    fx = TrainFixture("train_step_blender")
    r = _renderer(fx.cfg, fx.p_coarse, fx.p_fine)
    rays = fx.rays[:64].to(DEV)
    out = r.render_train(rays, None, want_z_fine=True)
    z_f = out["z_fine"].cpu()
    g = torch.Generator().manual_seed(8)
    gd_c, go_c = torch.randn(64, generator=g), torch.randn(64, generator=g)
    gr_f, gd_f = torch.randn(64, 3, generator=g), torch.randn(64, generator=g)
    gc, gf = r.backward(rays, None, {"coarse_comp_rgbs": torch.zeros(64, 3, device=DEV), "coarse_depth": gd_c.to(DEV),
                                      "coarse_opacity": go_c.to(DEV), "fine_comp_rgbs": gr_f.to(DEV), "fine_depth": gd_f.to(DEV)})
    pc = {k: v.double().requires_grad_(True) for k, v in fx.p_coarse.items()}
    pf = {k: v.double().requires_grad_(True) for k, v in fx.p_fine.items()}
    o = O.forward_rays(pc, pf, fx.rays[:64].double(), fx.cfg, None, z_fine_override=z_f.double())
    L = (o["coarse_depth"] * gd_c.double()).sum() + (o["coarse_opacity"] * go_c.double()).sum() + \
        (o["fine_comp_rgbs"] * gr_f.double()).sum() + (o["fine_depth"] * gd_f.double()).sum()
    L.backward()
    for flat, p in ((gc, pc), (gf, pf)):
        ref = torch.cat([v.grad.reshape(-1) for v in p.values()])
        err = _rel(flat, ref)
        _report(test="depth_opacity_grads", rel_l2=err)
        assert err < 5e-2, err
    r.close()


def test_depth_or_opacity_only_losses_reach_the_net():
    """a loss that touches only depth / opacity of a net (depth regularisers) used to get all-zero parameter
    gradients because nsr_backward skipped a net without a colour gradient."""
    fx = TrainFixture("train_step_blender")
    r = _renderer(fx.cfg, fx.p_coarse, fx.p_fine)
    rays = fx.rays[:64].to(DEV)
    z_f = r.render_train(rays, None, want_z_fine=True)["z_fine"].cpu()
    g = torch.Generator().manual_seed(9)
    gd_c, go_f = torch.randn(64, generator=g), torch.randn(64, generator=g)
    gc, gf = r.backward(rays, None, {"coarse_depth": gd_c.to(DEV), "fine_opacity": go_f.to(DEV)})
    pc = {k: v.double().requires_grad_(True) for k, v in fx.p_coarse.items()}
    pf = {k: v.double().requires_grad_(True) for k, v in fx.p_fine.items()}
    o = O.forward_rays(pc, pf, fx.rays[:64].double(), fx.cfg, None, z_fine_override=z_f.double())
    ((o["coarse_depth"] * gd_c.double()).sum() + (o["fine_opacity"] * go_f.double()).sum()).backward()
    for flat, p in ((gc, pc), (gf, pf)):
        want = torch.cat([v.grad.reshape(-1) for v in p.values()])
        assert float(flat.abs().max()) > 0.0
        cos = float(torch.nn.functional.cosine_similarity(flat.double().cpu()[None], want[None]))
        assert cos > 0.999, cos
    r.close()


def test_trainer_trajectory_against_oracle():
    """Five full iterations (forward, loss, backward, Adam, re-pack) on the same draws as the oracle's
    optimize_parameters: the loss trajectories must stay together."""
    from nerf_sr_b200 import Trainer
    fx = TrainFixture("train_step_blender")
    r = _renderer(fx.cfg, fx.p_coarse, fx.p_fine)
    tr = Trainer(r, fx.p_coarse, fx.p_fine, lr=fx.tcfg.lr, downscale=fx.s)
    state = T.TrainState(fx.p_coarse, fx.p_fine)
    g = torch.Generator().manual_seed(123)
    rays, tgt = fx.rays.to(DEV), fx.target.to(DEV)
    mine, ref = [], []
    for it in range(5):
        rng = fx.rng[it] if it < 2 else O.RenderRng.draw(fx.rays.shape[0], fx.cfg, g)
        m = tr.optimize_parameters(rays, tgt, rng_dict(rng))
        losses, _ = T.optimize_parameters(state, fx.rays, fx.target, fx.cfg, fx.tcfg, rng, fx.s)
        mine.append(float(m[0] + m[2])); ref.append(float(losses["tot"]))
    _report(test="trajectory", mine=mine, oracle=ref)
    assert mine[0] == pytest.approx(fx.meta["steps"][0]["tot"], rel=2e-3)
    for a, b in zip(mine, ref):
        assert a == pytest.approx(b, rel=0.08), (mine, ref)
    r.close()


def test_trainer_with_all_loss_terms_tracks_the_reference():
    """Two full iterations with every loss term on (variance, depth variance, SISR target, norm clipping, sigma noise,
    4x4 SS) on the draws of tests/golden/train_step_sr_var_s4.npz: the totals must follow the numbers the unmodified
    reference produced (oracle/make_golden_train.py)."""
    from nerf_sr_b200 import Trainer
    fx = TrainFixture("train_step_sr_var_s4")
    r = _renderer(fx.cfg, fx.p_coarse, fx.p_fine)
    tr = Trainer(r, fx.p_coarse, fx.p_fine, lr=fx.tcfg.lr, grad_clip_val=fx.tcfg.grad_clip_val, grad_clip_type=fx.tcfg.grad_clip_type,
                 **_trainer_kwargs(fx))
    rays, tgt, tsr = fx.rays.to(DEV), fx.target.to(DEV), fx.target_sr.to(DEV)
    tots = []
    for it in range(2):
        tr.optimize_parameters(rays, tgt, rng_dict(fx.rng[it]), target_sr=tsr, far=float(fx.rays[0, 7]))
        tots.append(float(tr.last_terms[0, 5] + tr.last_terms[1, 5]))
    want = [st["tot"] for st in fx.meta["steps"]]
    _report(test="trajectory_all_terms", mine=tots, reference=want)
    # step 0: same weights; only the fine sample positions differ at fp32 round-off.  step 1 is after one
    # Adam update, whose first step has magnitude lr regardless of gradient scale (sign-like), hence the wider band
    assert tots[0] == pytest.approx(want[0], rel=1e-2), (tots, want)
    assert tots[1] == pytest.approx(want[1], rel=0.1), (tots, want)
    r.close()


def test_trainer_checkpoint_roundtrip(tmp_path):
    """save_networks after a step, load into a fresh Trainer: identical renders; a shape mismatch is refused."""
    from nerf_sr_b200 import NsrError, Trainer
    from nerf_sr_b200 import checkpoints as K
    fx = TrainFixture("train_step_blender")
    r = _renderer(fx.cfg, fx.p_coarse, fx.p_fine)
    tr = Trainer(r, fx.p_coarse, fx.p_fine, downscale=fx.s)
    rays, tgt = fx.rays.to(DEV), fx.target.to(DEV)
    tr.optimize_parameters(rays, tgt, rng_dict(fx.rng[0]), lr=K.LrSchedule().lr)
    want = {k: v.clone() for k, v in r.forward_rays(rays).items()}
    tr.save_networks(str(tmp_path), 3)
    assert K.latest_epoch(str(tmp_path)) == 3
    sd_c, sd_f = K.load_networks(str(tmp_path), 3)
    assert all(torch.equal(sd_c[k], tr.state_dict(0)[k].cpu()) for k in sd_c) and not torch.equal(sd_c["sigma.weight"], fx.p_coarse["sigma.weight"])
    r2 = _renderer(fx.cfg, fx.p_coarse, fx.p_fine)
    tr2 = Trainer(r2, fx.p_coarse, fx.p_fine, downscale=fx.s)
    tr2.load_networks(str(tmp_path), 3)
    got = r2.forward_rays(rays)
    for k in want:
        assert torch.equal(got[k], want[k]), k
    bad = dict(sd_c)
    bad["sigma.weight"] = torch.zeros(1, 128)
    K.save_networks(str(tmp_path), 4, bad, sd_f)
    with pytest.raises(NsrError):
        tr2.load_networks(str(tmp_path), 4)
    r.close(); r2.close()


def test_training_reduces_the_loss_on_a_fixed_batch():
    from nerf_sr_b200 import Trainer
    cfg = O.RenderConfig(white_bkgd=True)
    pc, pf = O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17)
    r = _renderer(cfg, pc, pf)
    tr = Trainer(r, pc, pf, lr=5e-4, downscale=2)
    n_lr = 257                                      # ragged tile counts in both passes
    rays = O.synthetic_rays(n_lr * 4, 12, "blender").to(DEV)
    tgt = torch.rand(n_lr, 3, generator=torch.Generator().manual_seed(3)).to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(0)
    hist = []
    for it in range(40):
        m = tr.optimize_parameters(rays, tgt, tr.draw_rng(rays.shape[0], gen))
        hist.append(float(m[0] + m[2]))
    _report(test="fixed_batch_training", first=hist[0], last=hist[-1], hist=hist)
    assert all(np.isfinite(hist))
    assert min(hist[-10:]) < hist[0], hist
    # and the oracle, started from the trained weights, sees the same loss (eval of the CUDA-trained nets)
    sd_c = {k: v.cpu() for k, v in tr.state_dict(0).items()}
    sd_f = {k: v.cpu() for k, v in tr.state_dict(1).items()}
    with torch.no_grad():
        o = O.forward_rays(sd_c, sd_f, rays.cpu(), cfg)
    out = r.forward_rays(rays)
    mx, viol = O.tolerance_violations(out["coarse_comp_rgbs"].cpu(), o["coarse_comp_rgbs"])
    assert viol == 0.0, mx
    r.close()


def test_train_api_errors():
    from nerf_sr_b200 import NsrError, Renderer
    cfg = O.RenderConfig()
    r = Renderer(cfg, torch.device(DEV), precision="fp32_simt")
    r.load_state_dict(0, O.make_mlp_params(cfg, 4)); r.load_state_dict(1, O.make_mlp_params(cfg, 17))
    with pytest.raises(NsrError) as e:
        r.render_train(O.synthetic_rays(8, 1, "blender").to(DEV))
    assert e.value.code == 2
    r.close()
    r = _renderer(cfg, O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17))
    with pytest.raises(NsrError):
        r.backward(O.synthetic_rays(8, 1, "blender").to(DEV), None, {"coarse_comp_rgbs": torch.zeros(8, 3, device=DEV),
                                                                        "coarse_weights": torch.zeros(8, 64, device=DEV)})
    r.close()


def test_patch_model_train_mode_on_a_reference_lookalike():
    """INTEGRATION.md step 3 with grad enabled: forward_rays returns autograd-connected tensors, so the reference's own
    loss code -- here colour MSE plus its optional variance and depth-variance terms (models/nerf_downX_model.py:
    332-353,374-378) -- and loss.backward() fill p.grad of netCoarse / netFine from the CUDA backward."""
    import torch.nn as nn
    from types import SimpleNamespace
    from nerf_sr_b200 import patch_model
    fx = TrainFixture("train_step_llff_clip")          # sigma noise 1.0, LLFF-like rays

    class Net(nn.Module):      # parameter names = models/networks.py:149-180
        def __init__(self, p):
            super().__init__()
            for i in range(8):
                w = p[f"xyz_encoding_{i+1}.0.weight"]
                setattr(self, f"xyz_encoding_{i+1}", nn.Sequential(nn.Linear(w.shape[1], w.shape[0]), nn.ReLU(True)))
            self.xyz_encoding_final = nn.Linear(256, 256)
            self.dir_encoding = nn.Sequential(nn.Linear(283, 128), nn.ReLU(True))
            self.sigma = nn.Linear(256, 1)
            self.rgb = nn.Sequential(nn.Linear(128, 3), nn.Sigmoid())
            self.load_state_dict(p)

    class NeRFDownXModel:
        pass

    m = NeRFDownXModel()
    m.opt = SimpleNamespace(**{**fx.cfg.__dict__, "skips": list(fx.cfg.skips)})
    m.device = torch.device(DEV)
    m.netCoarse = nn.DataParallel(Net(fx.p_coarse).to(DEV))
    m.netFine = nn.DataParallel(Net(fx.p_fine).to(DEV))
    m.randomized = True
    m.forward_rays = lambda rays: (_ for _ in ()).throw(AssertionError("reference path must not run"))
    patch_model(m, "bf16x3")
    rays, tgt = fx.rays.to(DEV), fx.target.to(DEV)
    n, s = rays.shape[0], fx.s

    def losses(out, target, far):
        rc, rf = out["coarse_comp_rgbs"], out["fine_comp_rgbs"]
        L = torch.nn.functional.mse_loss(rc.reshape(-1, s * s, 3).mean(1), target) + \
            torch.nn.functional.mse_loss(rf.reshape(-1, s * s, 3).mean(1), target)
        L = L + 0.01 * torch.sum(torch.var(rc.reshape(-1, s * s, 3), dim=1)) + 0.01 * torch.sum(torch.var(rf.reshape(-1, s * s, 3), dim=1))
        L = L + 0.01 * torch.sum(torch.var(out["fine_depth"].reshape(-1, s * s, 1) / far, dim=1))
        return L

    torch.manual_seed(7)
    out = m.forward_rays(rays)
    assert out["fine_comp_rgbs"].requires_grad and not out["fine_weights"].requires_grad
    losses(out, tgt, 1.0).backward()
    # the same draws, in the reference's order (models/utils.py:41, :210, :73, :210)
    torch.manual_seed(7)
    rng = {"u_coarse": torch.rand(n, 64, device=DEV), "noise_coarse": torch.randn(n, 64, device=DEV),
           "u_fine": torch.rand(n, 64, device=DEV), "noise_fine": torch.randn(n, 128, device=DEV)}
    z_f = m._nsr_renderer.render_train(rays, rng, want_z_fine=True)["z_fine"].cpu().double()
    pc = {k: v.double().requires_grad_(True) for k, v in fx.p_coarse.items()}
    pf = {k: v.double().requires_grad_(True) for k, v in fx.p_fine.items()}
    orng = O.RenderRng(*[rng[k].cpu().double() for k in ("u_coarse", "noise_coarse", "u_fine", "noise_fine")])
    o = O.forward_rays(pc, pf, fx.rays.double(), fx.cfg, orng, z_fine_override=z_f)
    losses(o, fx.target.double(), 1.0).backward()
    for net, ref in ((m.netCoarse, pc), (m.netFine, pf)):
        named = dict(net.module.named_parameters())
        got = torch.cat([named[k].grad.reshape(-1) for k in ref]).cpu()
        want = torch.cat([v.grad.reshape(-1) for v in ref.values()])
        cos = float(torch.nn.functional.cosine_similarity(got.double()[None], want[None]))
        err = _rel(got, want)
        _report(test="patch_model_train", rel_l2=err, cos=cos)
        assert cos > 0.9995 and err < 3e-2, (err, cos)
    # eval mode afterwards picks up changed parameters (version counters) and runs without autograd
    with torch.no_grad():
        m.randomized = False
        e1 = m.forward_rays(rays)
        m.netFine.module.sigma.bias.add_(0.5)
        e2 = m.forward_rays(rays)
    assert not e1["fine_comp_rgbs"].requires_grad
    assert not torch.equal(e1["fine_opacity"], e2["fine_opacity"])
    m._nsr_renderer.close()


def test_two_forwards_in_flight_and_stash_guard():
    """The autograd bridge keeps a private stash per forward, so two forward_rays calls (e.g. main rays and reference-view
    rays, models/nerf_downX_model.py:318-324) can precede their backwards; and nsr_backward refuses a workspace that was
    filled for another batch size or never filled."""
    from nerf_sr_b200 import NsrError, RenderFunction
    from nerf_sr_b200.training import OUT_KEYS
    fx = TrainFixture("train_step_blender")
    r = _renderer(fx.cfg, fx.p_coarse, fx.p_fine)
    names = list(fx.p_coarse)
    pcs = [torch.nn.Parameter(fx.p_coarse[n].clone().to(DEV)) for n in names]
    pfs = [torch.nn.Parameter(fx.p_fine[n].clone().to(DEV)) for n in names]
    ra, rb = fx.rays[:128].to(DEV), fx.rays[128:192].to(DEV)

    def grads(order):
        for p in pcs + pfs:
            p.grad = None
        oa = dict(zip(OUT_KEYS, RenderFunction.apply(r, ra, None, len(pcs), None, *pcs, *pfs)))
        ob = dict(zip(OUT_KEYS, RenderFunction.apply(r, rb, None, len(pcs), None, *pcs, *pfs)))
        la, lb = oa["fine_comp_rgbs"].square().mean() + oa["coarse_comp_rgbs"].mean(), ob["fine_comp_rgbs"].mean()
        if order == "joint":
            (la + lb).backward()
        else:
            lb.backward(); la.backward()
        return torch.cat([p.grad.reshape(-1) for p in pcs + pfs]).clone()
    g1, g2 = grads("joint"), grads("separate")
    assert torch.isfinite(g1).all() and _rel(g1, g2) < 1e-6
    out = r.render_train(ra, None)
    with pytest.raises(NsrError):
        r.backward(rb, None, {"coarse_comp_rgbs": torch.zeros(64, 3, device=DEV)})       # stash holds 128 rays
    with pytest.raises(NsrError):
        r.backward(ra, None, {"coarse_comp_rgbs": torch.zeros(128, 3, device=DEV)}, ws=r.new_train_workspace(128))
    r.close()


@pytest.mark.parametrize("variant", ["softplus_gamma_lindisp", "color_none", "vanilla_viewdir_column"])
def test_gradients_option_variants(variant):
    """The backward's activation branches (rendering.py:70-73 softplus, nerf_downX_model.py:271 gamma, networks.py:173-176
    colour activation none) and the vanilla model's view-direction column, against the oracle's autograd in fp64."""
    kw = {"softplus_gamma_lindisp": dict(sigma_activation="softplus", gamma_correct=True, lindisp=True),
          "color_none": dict(color_activation="none", white_bkgd=True),
          "vanilla_viewdir_column": dict(white_bkgd=True, viewdir_offset=8)}[variant]
    cfg = O.RenderConfig(**kw)
    pc, pf = O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17)
    from nerf_sr_b200 import Renderer
    r = Renderer(cfg, torch.device(DEV), precision="bf16x3", viewdir_offset=cfg.viewdir_offset)
    r.load_state_dict(0, pc); r.load_state_dict(1, pf)
    n = 96
    rays = O.synthetic_rays(n, 21, "blender")
    if cfg.viewdir_offset == 8:
        vd = torch.nn.functional.normalize(torch.randn(n, 3, generator=torch.Generator().manual_seed(5)), dim=-1)
        rays = torch.cat([rays, vd], 1).contiguous()
    g = torch.Generator().manual_seed(9)
    g_c, g_f = torch.randn(n, 3, generator=g) / n, torch.randn(n, 3, generator=g) / n
    dev_rays = rays.to(DEV)
    z_f = r.render_train(dev_rays, None, want_z_fine=True)["z_fine"].cpu()
    gc, gf = r.backward(dev_rays, None, {"coarse_comp_rgbs": g_c.to(DEV), "fine_comp_rgbs": g_f.to(DEV)})
    p64c = {k: v.double().requires_grad_(True) for k, v in pc.items()}
    p64f = {k: v.double().requires_grad_(True) for k, v in pf.items()}
    o = O.forward_rays(p64c, p64f, rays.double(), cfg, None, z_fine_override=z_f.double())
    ((o["coarse_comp_rgbs"] * g_c.double()).sum() + (o["fine_comp_rgbs"] * g_f.double()).sum()).backward()
    for flat, ref in ((gc, p64c), (gf, p64f)):
        want = torch.cat([v.grad.reshape(-1) for v in ref.values()])
        assert torch.isfinite(flat).all()
        err = _rel(flat, want)
        cos = float(torch.nn.functional.cosine_similarity(flat.double().cpu()[None], want[None]))
        _report(test="grad_variants", variant=variant, rel_l2=err, cos=cos)
        assert cos > 0.9995 and err < 3e-2, (variant, err, cos)
    r.close()


def test_fused_dx_chain_equals_layerwise_kernels():
    """k_tg_dxchain (dZ stays in TMEM between layers) against the layer-by-layer k_tg_dx launches (debug flag 2):
    same MMAs in the same order, so the gradients must agree to rounding; ragged tile counts in both passes."""
    from nerf_sr_b200 import Trainer
    cfg = O.RenderConfig(white_bkgd=True, noise_std=0.5)
    pc, pf = O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17)
    r = _renderer(cfg, pc, pf)
    tr = Trainer(r, pc, pf, downscale=2)
    n_lr = 151 * 3 + 1                                   # > 148 coarse tiles per CTA round, odd ray count per tile pair
    rays = O.synthetic_rays(n_lr * 4, 33, "blender").to(DEV)
    tgt = torch.rand(n_lr, 3, generator=torch.Generator().manual_seed(4)).to(DEV)
    rng = tr.draw_rng(rays.shape[0], torch.Generator(device=DEV).manual_seed(1))
    gc1, gf1 = tr.forward_backward(rays, tgt, rng)
    r.lib.nsr_debug_set_flags(r._h, 2)
    gc2, gf2 = tr.forward_backward(rays, tgt, rng)
    r.lib.nsr_debug_set_flags(r._h, 0)
    torch.cuda.synchronize()
    e1, e2 = _rel(gc1, gc2), _rel(gf1, gf2)
    _report(test="chain_vs_layerwise", coarse=e1, fine=e2)
    assert torch.isfinite(gc1).all() and torch.isfinite(gf1).all()
    assert e1 < 1e-6 and e2 < 1e-6, (e1, e2)
    r.close()


def test_dw_stream_fan_out_is_bit_identical_and_stream_ordered():
    """nsr_backward spreads a net's independent dW GEMMs over three streams (fork / join by events).  The result must be
    bit-identical to the single-stream order (debug flag 4) -- same kernels, fixed-order reduction -- also when the call
    is issued on a non-default stream with work queued right behind it, repeatedly (join really orders the consumers)."""
    from nerf_sr_b200 import Trainer
    cfg = O.RenderConfig(noise_std=1.0)
    pc, pf = O.make_mlp_params(cfg, 21), O.make_mlp_params(cfg, 8)
    r = _renderer(cfg, pc, pf)
    tr = Trainer(r, pc, pf, downscale=2)
    n_lr = 300
    rays = O.synthetic_rays(n_lr * 4, 9, "llff").to(DEV)
    tgt = torch.rand(n_lr, 3, generator=torch.Generator().manual_seed(4)).to(DEV)
    rng = tr.draw_rng(rays.shape[0], torch.Generator(device=DEV).manual_seed(1))
    r.lib.nsr_debug_set_flags(r._h, 4)
    gc0, gf0 = tr.forward_backward(rays, tgt, rng)
    gc0, gf0 = gc0.clone(), gf0.clone()
    r.lib.nsr_debug_set_flags(r._h, 0)
    torch.cuda.synchronize()
    side = torch.cuda.Stream(torch.device(DEV))
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(5):
            gc, gf = tr.forward_backward(rays, tgt, rng)
            sc, sf = gc.clone(), gf.clone()             # consumers queued immediately behind the join
            assert torch.equal(sc, gc0) and torch.equal(sf, gf0)
    side.synchronize()
    r.close()
