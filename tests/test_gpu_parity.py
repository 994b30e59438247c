"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every check goes through the C ABI
(nerf_sr_b200.Renderer -> ctypes -> libnsr_b200.so) and compares with the committed golden
fixtures (outputs of the unmodified reference) or with the oracle on seeded inputs.

Tolerance (BASELINE.json north_star): |a-b| <= 1e-4 + 1e-3*|b|, fp32.
Protocol: (i) coarse stage direct; (ii) fine stage teacher-forced on the
reference's fine z-values; (iii) end-to-end fine as a violation fraction bounded by the
reference's own fp32-vs-fp64 floor (+ margin), because fine sample positions are an
ill-conditioned function of coarse weights; (iv) resampler unit parity."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, golden_names
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu

PRECISIONS = ["fp32_simt", "bf16x3", "fp16x3"]
COARSE_KEYS = ("coarse_comp_rgbs", "coarse_depth", "coarse_opacity", "coarse_weights")
FINE_MAP = (("comp_rgbs", "fine_comp_rgbs"), ("depth", "fine_depth"), ("opacity", "fine_opacity"),
            ("weights", "fine_weights"))


def _renderer(fx, prec):
    from nerf_sr_b200 import NsrError, Renderer
    try:
        r = Renderer(fx.cfg, torch.device("cuda:0"), precision=prec, viewdir_offset=fx.cfg.viewdir_offset)
    except NsrError as e:
        if e.code == 2:
            pytest.skip(f"{prec} does not support this option set: {e}")
        raise
    r.load_state_dict(0, fx.p_coarse)
    r.load_state_dict(1, fx.p_fine)
    return r


def _rng(fx):
    if fx.rng is None:
        return None
    return {k: getattr(fx.rng, k) for k in ("u_coarse", "noise_coarse", "u_fine", "noise_fine")
            if getattr(fx.rng, k) is not None}


def _dev(t):
    return None if t is None else t.cuda()


def _report(**rec):
    import json
    import os
    from conftest import ROOT
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "r02_parity.jsonl"), "a") as fh:
        fh.write(json.dumps(rec) + "\n")


_FP64 = {}


def _fp64_reference(fx):
    """The oracle in fp64 on cuda:0 for a golden fixture (same weights, rays and train-mode draws): the 'exact' answer
    both the reference's fp32 output (the fixture) and ours are measured against."""
    if fx.name not in _FP64:
        dev = torch.device("cuda:0")
        d = lambda t: None if t is None else t.to(dev).double()
        rng = None
        if fx.rng is not None:
            rng = O.RenderRng(d(fx.rng.u_coarse), d(fx.rng.noise_coarse), d(fx.rng.u_fine), d(fx.rng.noise_fine))
        with torch.no_grad():
            out = O.forward_rays({k: d(v) for k, v in fx.p_coarse.items()}, {k: d(v) for k, v in fx.p_fine.items()},
                                 d(fx.rays), fx.cfg, rng)
        _FP64[fx.name] = {k: v.cpu() for k, v in out.items()}
    return _FP64[fx.name]


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("name", golden_names())
def test_forward_rays_against_reference_golden(name, prec, load_fixture):
    fx = load_fixture(name)
    r = _renderer(fx, prec)
    out = r.forward_rays(fx.rays.cuda(), _rng(fx), want_z_fine=True)
    torch.cuda.synchronize()
    assert set(fx.out) <= set(out)
    for k, ref in fx.out.items():
        assert out[k].shape == ref.shape and out[k].dtype == torch.float32, k
    # (i) coarse stage: direct
    for k in COARSE_KEYS:
        mx, viol = O.tolerance_violations(out[k].cpu(), fx.out[k])
        assert viol == 0.0, (k, mx, viol)
    if fx.cfg.N_importance == 0:
        assert "fine_comp_rgbs" not in out
        return
    # (iii) end to end: the rule of conftest.e2e_bounds (floor = the reference's OWN fp32-vs-fp64 disagreement on this fixture,
    # recomputed here with the oracle in fp64 on the GPU; margin per precision; binomial 3 sigma of a count on n_rays rays)
    from conftest import e2e_bounds
    ref64 = _fp64_reference(fx)
    for k in ("fine_comp_rgbs", "fine_depth", "fine_opacity", "fine_weights"):
        got = out[k].cpu()
        mx, viol = O.tolerance_violations(got, fx.out[k])
        _, floor = O.tolerance_violations(fx.out[k], ref64[k])
        _, v64 = O.tolerance_violations(got, ref64[k])
        _report(test="e2e_fine", fixture=name, prec=prec, key=k, max_abs=mx, viol_vs_ref32=viol, viol_vs_fp64=v64, floor=floor,
                floor_in_fixture=fx.meta["fp64_floor"][k]["viol"])
        b64, b32 = e2e_bounds(floor, fx.rays.shape[0], prec)
        assert v64 <= b64, (k, v64, floor, b64)
        assert viol <= b32, (k, mx, viol, floor, b32)
        assert torch.isfinite(out[k]).all()
    r.close()


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("name", golden_names())
def test_teacher_forced_passes_and_mlp_output(name, prec, load_fixture):
    """(ii): each network on the reference's own z-values, plus the raw VanillaMLP output."""
    fx = load_fixture(name)
    r = _renderer(fx, prec)
    rays = fx.rays.cuda()
    nz = _dev(fx.rng.noise_coarse) if fx.rng is not None else None
    pc = r.render_pass(0, rays, fx.z_coarse.cuda(), nz, want_raw=True)
    for kl, kr in (("comp_rgbs", "coarse_comp_rgbs"), ("depth", "coarse_depth"), ("opacity", "coarse_opacity"),
                   ("weights", "coarse_weights")):
        mx, viol = O.tolerance_violations(pc[kl].cpu(), fx.out[kr])
        assert viol == 0.0, (kl, mx, viol)
    mx, viol = O.tolerance_violations(pc["raw"].cpu(), fx.raw_coarse, rtol=1e-3, atol=1e-4)     # north_star tolerance
    _report(test="raw_mlp", fixture=name, prec=prec, net="coarse", max_abs=mx, viol=viol)
    assert viol == 0.0, ("raw_coarse", mx, viol)
    if fx.cfg.N_importance > 0:
        nzf = _dev(fx.rng.noise_fine) if fx.rng is not None else None
        pf = r.render_pass(1, rays, fx.z_fine.cuda(), nzf, want_raw=True)
        for kl, kr in FINE_MAP:
            mx, viol = O.tolerance_violations(pf[kl].cpu(), fx.out[kr])
            assert viol == 0.0, (kl, mx, viol)
        mx, viol = O.tolerance_violations(pf["raw"].cpu(), fx.raw_fine, rtol=1e-3, atol=1e-4)
        _report(test="raw_mlp", fixture=name, prec=prec, net="fine", max_abs=mx, viol=viol)
        assert viol == 0.0, ("raw_fine", mx, viol)
    r.close()


@pytest.mark.parametrize("name", ["eval_blender", "eval_llff", "train_blender", "train_llff_noise", "opt_small_net"])
def test_sampling_seams(name, load_fixture):
    """(iv) sample_along_rays / resample_along_rays z-values against the reference's."""
    fx = load_fixture(name)
    r = _renderer(fx, "fp32_simt")
    u = _dev(fx.rng.u_coarse) if fx.rng is not None else None
    z = r.sample_along_rays(fx.rays.cuda(), u)
    assert torch.equal(z.cpu(), fx.z_coarse), float((z.cpu() - fx.z_coarse).abs().max())   # bit exact
    uf = _dev(fx.rng.u_fine) if fx.rng is not None else None
    zf = r.resample_along_rays(fx.z_coarse.cuda(), fx.out["coarse_weights"].cuda(), uf).cpu()
    assert (zf[:, 1:] >= zf[:, :-1]).all()                  # sorted
    mx, viol = O.tolerance_violations(zf, fx.z_fine)
    # flat CDF segments amplify 1-ulp differences of the pdf normalisation: tolerance, not bit-exact
    assert viol == 0.0, (mx, viol)
    assert float((zf - fx.z_fine).abs().median()) < 1e-6
    r.close()


def test_posenc_and_box_average(load_fixture):
    fx = load_fixture("eval_blender")
    r = _renderer(fx, "fp32_simt")
    x = torch.randn(1000, 3, generator=torch.Generator().manual_seed(0)) * 3
    for deg in (10, 4):
        e = r.posenc(x.cuda(), deg).cpu()
        ref = O.posenc(x, deg)
        assert e.shape == ref.shape
        assert torch.equal(e[:, :3], x)
        # sin / cos of arguments up to 2^9 |x| ~ 5000 rad.  The yardstick is the EXACT function of the fp32-rounded argument
        # f * x (the product is what every implementation agrees on bit for bit): host libm / SLEEF builds differ from one
        # another by up to 1.5e-4 out there (seen between two GPU boxes of this pool), CUDA's sinf / cosf stay within 2 ulp.
        exact = [x.double()]
        for f in O.frequency_bands(deg):
            arg = (f * x).double()                         # fp32 product, then exact evaluation
            exact += [torch.sin(arg), torch.cos(arg)]
        exact = torch.cat(exact, -1)
        assert float((e.double() - exact).abs().max()) < 5e-7
        assert float((e - ref).abs().max()) < 1e-3          # and the host's fp32 evaluation is in the same place
    v = torch.rand(4 * 50, 3, generator=torch.Generator().manual_seed(1))
    assert torch.allclose(r.box_average(v.cuda(), 2).cpu(), O.box_average(v, 2), rtol=0, atol=1e-7)
    d = torch.rand(16 * 10, generator=torch.Generator().manual_seed(2))
    assert torch.allclose(r.box_average(d.cuda(), 4).cpu(), O.box_average(d, 4), rtol=0, atol=1e-7)
    r.close()


def test_generate_rays_matches_reference_dataset_path(load_fixture):
    import os
    fx = load_fixture("eval_blender")
    r = _renderer(fx, "fp32_simt")
    z = np.load(os.path.join(GOLDEN_DIR, "raygen.npz"))
    for tag in ("blender", "llff"):
        H, W, s, focal, ndc, near, far = z[f"{tag}_params"]
        rays = r.generate_rays(z[f"{tag}_c2w"], int(H), int(W), float(focal), int(s), bool(ndc), float(near), float(far)).cpu()
        ref = torch.from_numpy(z[f"{tag}_rays"])
        assert rays.shape == ref.shape
        assert torch.allclose(rays, ref, rtol=1e-5, atol=1e-5), float((rays - ref).abs().max())
    r.close()


@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3"])
def test_full_size_frame_properties(prec, load_fixture):
    """BASELINE configs[1] size (160 000 rays): size-independent properties + agreement of the
    tensor-core path with the fp32 CUDA-core path + determinism + ragged tail."""
    fx = load_fixture("eval_blender")
    cfg = fx.cfg
    rays = O.synthetic_rays(160000 + 3, 11, "blender").cuda()      # odd count: ragged last tile
    r = _renderer(fx, prec)
    a = r.forward_rays(rays)
    b = r.forward_rays(rays)
    torch.cuda.synchronize()
    for k in a:
        assert torch.equal(a[k], b[k]), f"{k} not deterministic"
        assert torch.isfinite(a[k]).all(), k
    for p in ("coarse", "fine"):
        w, op = a[f"{p}_weights"], a[f"{p}_opacity"]
        assert float(w.min()) >= 0.0 and float(op.max()) <= 1.0 + 1e-5
        assert torch.allclose(w.sum(-1), op, rtol=1e-5, atol=1e-5)            # opacity == sum of weights
        rgb = a[f"{p}_comp_rgbs"]
        assert float(rgb.min()) >= -1e-5 and float(rgb.max()) <= 1.0 + 1e-4    # white bkgd + sigmoid colours
        assert float(a[f"{p}_depth"].max()) <= 6.0 + 1e-3
    # the coarse stage must agree with the fp32 CUDA-core path within tolerance everywhere
    s = _renderer(fx, "fp32_simt")
    c = s.forward_rays(rays[:20000])
    for k in COARSE_KEYS:
        mx, viol = O.tolerance_violations(a[k][:20000].cpu(), c[k].cpu())
        assert viol == 0.0, (k, mx, viol)
    # a chunked call equals the single call (no cross-ray state; chunk_batch equivalence, utils.py:130-152)
    d = r.forward_rays(rays[4096:8192])
    for k in d:
        assert torch.equal(d[k], a[k][4096:8192]), k
    r.close(); s.close()


def test_host_frame_call_equals_device_path(load_fixture):
    fx = load_fixture("eval_blender")
    r = _renderer(fx, "bf16x3")
    rays = O.synthetic_rays(4 * 40000 + 4 * 7, 5, "blender")
    rgb, depth = r.render_frame_host(rays, 2)
    out = r.forward_rays(rays.cuda(), want_weights=False)
    assert torch.equal(rgb, r.box_average(out["fine_comp_rgbs"], 2).cpu())
    assert torch.equal(depth, r.box_average(out["fine_depth"], 2).cpu().squeeze(-1))
    r.close()


def test_errors_are_reported_not_fatal(load_fixture):
    from nerf_sr_b200 import NsrError, Renderer
    fx = load_fixture("eval_blender")
    r = Renderer(fx.cfg, torch.device("cuda:0"), precision="bf16x3")
    with pytest.raises(NsrError) as ei:                    # weights not packed yet
        r.forward_rays(fx.rays.cuda())
    assert ei.value.code == 3
    bad = O.RenderConfig(W=100)
    with pytest.raises(NsrError):
        Renderer(bad, torch.device("cuda:0"), precision="fp32_simt")
    with pytest.raises(NsrError):                          # tensor-core path refuses, never silently differs
        Renderer(O.RenderConfig(D=4), torch.device("cuda:0"), precision="bf16x3")
    r.close()


def test_patch_model_rebinds_forward_rays_on_a_reference_lookalike(load_fixture):
    """INTEGRATION.md step 3 on an object with the reference model's attribute surface
    (opt, device, netCoarse/netFine with the reference's state_dict names, randomized)."""
    import torch.nn as nn
    from types import SimpleNamespace
    from nerf_sr_b200 import patch_model
    fx = load_fixture("eval_blender")

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

    class NeRFDownXModel:      # name matters: viewdir column 3
        pass

    m = NeRFDownXModel()
    m.opt = SimpleNamespace(**{**fx.cfg.__dict__, "skips": list(fx.cfg.skips)})
    m.device = torch.device("cuda:0")
    m.netCoarse = nn.DataParallel(Net(fx.p_coarse).cuda())     # 'module.' unwrap path (base_model.py:193-194)
    m.netFine = nn.DataParallel(Net(fx.p_fine).cuda())
    m.randomized = False
    m.forward_rays = lambda rays: (_ for _ in ()).throw(AssertionError("reference path must not run in no_grad"))
    patch_model(m, "bf16x3")
    with torch.no_grad():
        out = m.forward_rays(fx.rays.cuda())
    for k in COARSE_KEYS:
        mx, viol = O.tolerance_violations(out[k].cpu(), fx.out[k])
        assert viol == 0.0, (k, mx)
    assert float(m.near[0]) == 2.0 and float(m.far[0]) == 6.0
    # weights are re-packed when parameters change (optimizer step / load_networks)
    with torch.no_grad():
        m.netCoarse.module.sigma.bias.add_(0.25)
        out2 = m.forward_rays(fx.rays.cuda())
    assert not torch.equal(out2["coarse_opacity"], out["coarse_opacity"])
    m._nsr_renderer.close()


def test_host_pipeline_is_ordered_behind_pack_weights(load_fixture):
    """nsr_pack_weights is asynchronous on the caller's stream while nsr_render_host /
    nsr_render_pose_host run on library-owned streams; a pack immediately followed by a host render must see the NEW
    weights, with no host synchronisation in between."""
    from nerf_sr_b200 import Renderer
    fx = load_fixture("eval_blender")
    dev = torch.device("cuda:0")
    rays = O.synthetic_rays(4 * 5000, 5, "blender")
    pose = torch.tensor([[1.0, 0, 0, 0.1], [0, 1, 0, -0.2], [0, 0, 1, 4.0]])
    other_c, other_f = O.make_mlp_params(fx.cfg, 31), O.make_mlp_params(fx.cfg, 34)
    want = Renderer(fx.cfg, dev, precision="bf16x3")
    want.load_state_dict(0, other_c)
    want.load_state_dict(1, other_f)
    torch.cuda.synchronize()
    rgb_want, depth_want = want.render_frame_host(rays, 2)
    prgb_want, _ = want.render_pose_host(pose, 64, 64, 80.0, 2)
    r = _renderer(fx, "bf16x3")
    r.render_frame_host(rays, 2)                     # creates the library streams, old weights
    busy = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        for _ in range(8):
            busy.fill_(1)                            # keep the caller's stream busy so that the pack is still queued
        r.load_state_dict(0, other_c)
        r.load_state_dict(1, other_f)
        rgb, depth = r.render_frame_host(rays, 2)
        assert torch.equal(rgb, rgb_want) and torch.equal(depth, depth_want)
        for _ in range(8):
            busy.fill_(2)
        r.load_state_dict(0, fx.p_coarse)
        r.load_state_dict(1, fx.p_fine)
        for _ in range(8):
            busy.fill_(3)
        r.load_state_dict(0, other_c)
        r.load_state_dict(1, other_f)
        prgb, _ = r.render_pose_host(pose, 64, 64, 80.0, 2)
        assert torch.equal(prgb, prgb_want)
    r.close()
    want.close()


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("n", [1, 2, 3, 129])
def test_tiny_and_ragged_batches(n, prec, load_fixture):
    """Edge sizes: a single ray, an odd count (half-filled coarse tile), more tiles than one."""
    fx = load_fixture("eval_blender")
    r = _renderer(fx, prec)
    full = r.forward_rays(fx.rays.cuda())
    part = r.forward_rays(fx.rays[:n].cuda())
    torch.cuda.synchronize()
    for k in part:
        assert part[k].shape[0] == n
        assert torch.equal(part[k], full[k][:n]), k
    empty = r.forward_rays(fx.rays[:0].cuda())
    assert all(v.shape[0] == 0 for v in empty.values())
    r.close()


def test_lr_metrics_match_reference_losses(load_fixture):
    """Row f-2: box average + ColorMSELoss + PSNR (models/criterions.py) in one call."""
    fx = load_fixture("eval_blender")
    r = _renderer(fx, "fp32_simt")
    g = torch.Generator().manual_seed(3)
    for s, n_lr in ((2, 5000), (4, 777), (1, 333)):
        hr = torch.rand(n_lr * s * s, 3, generator=g)
        tgt = torch.rand(n_lr, 3, generator=g)
        lr, m = r.lr_metrics(hr.cuda(), tgt.cuda(), s)
        ref_lr, ref_mse, ref_psnr = O.lr_metrics(hr, tgt, s)
        assert torch.allclose(lr.cpu(), ref_lr, rtol=0, atol=5e-7)     # fp32 summation order of the 16-term mean
        assert abs(float(m[0]) - float(ref_mse)) <= 1e-6 * float(ref_mse)
        assert abs(float(m[1]) - float(ref_psnr)) <= 1e-5 * abs(float(ref_psnr)) + 1e-5
    r.close()


@pytest.mark.parametrize("prec", ["bf16x3", "fp32_simt"])
def test_single_pass_128_samples_against_oracle(prec):
    """(N_coarse, N_importance) = (128, 0): the other tile shape of the tensor-core path (1 ray / tile
    with on-the-fly z sampling).  No golden fixture: compared with the pinned oracle on seeded inputs."""
    from nerf_sr_b200 import Renderer
    cfg = O.RenderConfig(N_coarse=128, N_importance=0, white_bkgd=True)
    pc = O.make_mlp_params(cfg, 4)
    rays = O.synthetic_rays(96, 9, "blender")
    with torch.no_grad():
        ref = O.forward_rays(pc, pc, rays, cfg)
    r = Renderer(cfg, torch.device("cuda:0"), precision=prec)
    r.load_state_dict(0, pc)
    r.load_state_dict(1, pc)
    out = r.forward_rays(rays.cuda())
    assert set(out) == set(ref)
    for k in ref:
        mx, viol = O.tolerance_violations(out[k].cpu(), ref[k])
        assert viol == 0.0, (k, mx)
    r.close()


def test_render_pose_host_equals_rays_path(load_fixture):
    """Row f-4 core: pose -> device ray generation -> render -> box average -> host, against the
    explicit path (generate_rays + forward_rays + box_average)."""
    fx = load_fixture("eval_blender")
    r = _renderer(fx, "bf16x3")
    z = np.load(__import__("os").path.join(GOLDEN_DIR, "raygen.npz"))
    c2w = z["blender_c2w"]
    H, W, s, focal = 96, 128, 2, 150.0
    rgb, depth = r.render_pose_host(c2w, H, W, focal, s, False, 2.0, 6.0)
    rays = r.generate_rays(c2w, H, W, focal, s, False, 2.0, 6.0)
    out = r.forward_rays(rays, want_weights=False)
    assert torch.equal(rgb, r.box_average(out["fine_comp_rgbs"], s).cpu())
    assert torch.equal(depth, r.box_average(out["fine_depth"], s).cpu().squeeze(-1))
    assert rgb.shape == ((H // s) * (W // s), 3) and torch.isfinite(rgb).all()
    r.close()


@pytest.mark.parametrize("prec", ["bf16x3"])
def test_tile_count_stress_no_deadlock_and_chunk_invariance(prec, load_fixture):
    """The tcgen05 kernel is a persistent, barrier-driven pipeline: sweep ray counts so that CTAs get
    0, 1, 2, 3, ... tiles each (148 SMs; 2 rays per coarse tile, 1 per fine tile), in eval and train
    mode.  A protocol error shows up as a trapped launch (in-kernel mbarrier watchdog), a wrong tail
    as a mismatch with the corresponding slice of one big batch."""
    fx = load_fixture("eval_blender")
    r = _renderer(fx, prec)
    n_max = 148 * 7 + 5
    rays = O.synthetic_rays(n_max, 21, "blender").cuda()
    g = torch.Generator().manual_seed(5)
    rng = {"u_coarse": torch.rand(n_max, 64, generator=g).cuda(), "u_fine": torch.rand(n_max, 64, generator=g).cuda()}
    full = r.forward_rays(rays)
    full_t = r.forward_rays(rays, rng)
    sizes = sorted({1, 2, 3, 5, 64, 127, 128, 129, 147, 148, 149, 150, 295, 296, 297, 298, 443, 444, 445, 591, 592, 593,
                    148 * 4 - 1, 148 * 4, 148 * 4 + 1, 148 * 6 + 1, 148 * 7 + 5})
    for n in sizes:
        out = r.forward_rays(rays[:n])
        for k in out:
            assert torch.equal(out[k], full[k][:n]), (n, k)
        sub = {k: v[:n] for k, v in rng.items()}
        out_t = r.forward_rays(rays[:n], sub)
        for k in out_t:
            assert torch.equal(out_t[k], full_t[k][:n]), (n, k, "train")
    torch.cuda.synchronize()
    r.close()


@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3"])
def test_against_oracle_run_on_the_gpu(prec, load_fixture):
    """the same ATen op sequence on cuda:0 (TF32 off) is the bit-closest oracle for GPU
    sin/cos/exp.  Coarse stage must hold the tolerance on a larger, unseen ray set; the end-to-end fine
    stage is reported against the same floor rule as the golden fixtures."""
    torch.backends.cuda.matmul.allow_tf32 = False
    fx = load_fixture("eval_llff")
    from conftest import e2e_bounds
    rays = O.synthetic_rays(32768, 77, "llff").cuda()
    pc = {k: v.cuda() for k, v in fx.p_coarse.items()}
    pf = {k: v.cuda() for k, v in fx.p_fine.items()}
    with torch.no_grad():
        ref = O.forward_rays(pc, pf, rays, fx.cfg)
    r = _renderer(fx, prec)
    out = r.forward_rays(rays)
    for k in COARSE_KEYS:
        mx, viol = O.tolerance_violations(out[k].cpu(), ref[k].cpu())
        assert viol == 0.0, (k, mx, viol)
    d = lambda p: {k: v.double() for k, v in p.items()}
    with torch.no_grad():
        ref64 = O.forward_rays(d(pc), d(pf), rays.double(), fx.cfg)
    for k in ("fine_comp_rgbs", "fine_depth", "fine_weights"):
        mx, viol = O.tolerance_violations(out[k].cpu(), ref[k].cpu())
        _, floor = O.tolerance_violations(ref[k].cpu(), ref64[k].cpu())
        _, v64 = O.tolerance_violations(out[k].cpu(), ref64[k].cpu())
        _report(test="e2e_fine_gpu_oracle", prec=prec, key=k, max_abs=mx, viol_vs_ref32=viol, viol_vs_fp64=v64, floor=floor)
        b64, b32 = e2e_bounds(floor, rays.shape[0], prec)
        assert v64 <= b64, (k, v64, floor, b64)
        assert viol <= b32, (k, mx, viol, floor, b32)
    r.close()


@pytest.mark.parametrize("W", [128, 64])
@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3"])
def test_narrower_nets_run_zero_padded_on_the_tensor_core_path(W, prec):
    """--W 128 and friends (models/networks.py:125): weights, biases and head weights are zero-padded to the kernel's 256 /
    128 features, which is exact (a padded activation is relu(0) = 0 and adds exact zeros).  Same acceptance rules as the
    256-wide net, against the oracle's ATen op sequence on the GPU; one launch per call proves it is the tcgen05 frame kernel
    and not the fp32 CUDA-core path."""
    from conftest import e2e_bounds
    from nerf_sr_b200 import Renderer
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = O.RenderConfig(W=W, white_bkgd=True)
    pc, pf = O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17)
    assert pc["xyz_encoding_2.0.weight"].shape == (W, W) and pc["dir_encoding.0.weight"].shape == (W // 2, W + 27)
    rays = O.synthetic_rays(4096, 60 + W, "blender").cuda()
    pcd, pfd = {k: v.cuda() for k, v in pc.items()}, {k: v.cuda() for k, v in pf.items()}
    ex = {}
    with torch.no_grad():
        ref = O.forward_rays(pcd, pfd, rays, cfg, extras=ex)
        d = lambda p: {k: v.double() for k, v in p.items()}
        ref64 = O.forward_rays(d(pcd), d(pfd), rays.double(), cfg)
    r = Renderer(cfg, torch.device("cuda:0"), precision=prec)
    r.load_state_dict(0, pc)
    r.load_state_dict(1, pf)
    l0 = r.launch_count
    out = r.forward_rays(rays)
    assert r.launch_count - l0 == 1
    for k in COARSE_KEYS:
        mx, viol = O.tolerance_violations(out[k], ref[k])
        assert viol == 0.0, (k, mx, viol)
    tf = r.render_pass(1, rays, ex["z_fine"], want_raw=True)
    for kl, kr in (("comp_rgbs", "fine_comp_rgbs"), ("depth", "fine_depth"), ("opacity", "fine_opacity"), ("weights", "fine_weights")):
        mx, viol = O.tolerance_violations(tf[kl], ref[kr])
        assert viol == 0.0, (kl, mx, viol)
    mx, viol = O.tolerance_violations(tf["raw"], ex["raw_fine"])
    assert viol == 0.0, ("raw", mx, viol)
    for k in ("fine_comp_rgbs", "fine_depth", "fine_weights"):
        _, viol = O.tolerance_violations(out[k], ref[k])
        _, floor = O.tolerance_violations(ref[k], ref64[k])
        _, v64 = O.tolerance_violations(out[k], ref64[k])
        _report(test="narrow_net", W=W, prec=prec, key=k, viol_vs_ref32=viol, viol_vs_fp64=v64, floor=floor)
        b64, b32 = e2e_bounds(floor, rays.shape[0], prec)
        assert v64 <= b64 and viol <= b32, (k, viol, v64, floor)
    with pytest.raises(Exception):                       # training stays with the 256-wide net
        from nerf_sr_b200 import Trainer
        Trainer(r, pc, pf, downscale=2).optimize_parameters(rays[:64], torch.rand(16, 3, device="cuda"), None)
    r.close()


def test_single_pass_bf16_fast_mode_is_close_but_not_parity_grade(load_fixture):
    """NSR_PREC_BF16_TC: one bf16 MMA per product.  Documented as NOT parity grade; it must
    still be a faithful render: high PSNR against the reference, finite, same shapes."""
    fx = load_fixture("eval_llff")
    r = _renderer(fx, "bf16")
    out = r.forward_rays(fx.rays.cuda())
    for k, ref in fx.out.items():
        assert out[k].shape == ref.shape and torch.isfinite(out[k]).all(), k
    mse = float(((out["coarse_comp_rgbs"].cpu() - fx.out["coarse_comp_rgbs"]) ** 2).mean())
    assert -10 * np.log10(mse + 1e-20) > 20.0     # random-init nets amplify bf16 rounding (~1e-1 max err)
    _, viol = O.tolerance_violations(out["coarse_comp_rgbs"].cpu(), fx.out["coarse_comp_rgbs"])
    assert viol > 0.0          # i.e. the 3-pass split is what buys parity
    r.close()


@pytest.mark.parametrize("opts,precs", [(dict(no_logscale=True), ["fp32_simt", "bf16x3"]),
                                         (dict(no_xyz=True), ["fp32_simt"]),
                                         (dict(no_xyz=True, no_logscale=True, deg_pos=6, deg_dir=2), ["fp32_simt"])])
def test_embedding_options_against_oracle(opts, precs):
    """--no_logscale / --no_xyz (models/embedding.py:17-18,39-42): compared with the pinned oracle on
    seeded inputs; the tensor-core path must refuse --no_xyz rather than differ silently."""
    from nerf_sr_b200 import NsrError, Renderer
    cfg = O.RenderConfig(white_bkgd=True, **opts)
    pc, pf = O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17)
    rays = O.synthetic_rays(160, 13, "blender")
    ex = {}
    with torch.no_grad():
        ref = O.forward_rays(pc, pf, rays, cfg, extras=ex)
    for prec in precs:
        r = Renderer(cfg, torch.device("cuda:0"), precision=prec)
        r.load_state_dict(0, pc)
        r.load_state_dict(1, pf)
        out = r.forward_rays(rays.cuda())
        for k in COARSE_KEYS:
            mx, viol = O.tolerance_violations(out[k].cpu(), ref[k])
            assert viol == 0.0, (prec, k, mx)
        p = r.render_pass(1, rays.cuda(), ex["z_fine"].cuda())
        for kl, kr in FINE_MAP:
            mx, viol = O.tolerance_violations(p[kl].cpu(), ref[kr])
            assert viol == 0.0, (prec, kl, mx)
        r.close()
    if opts.get("no_xyz"):
        with pytest.raises(NsrError) as ei:
            Renderer(cfg, torch.device("cuda:0"), precision="bf16x3")
        assert ei.value.code == 2


@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3"])
@pytest.mark.parametrize("n_coarse,n_importance", [(64, 128), (128, 64), (128, 128), (64, 192)])
def test_more_fine_samples_stay_on_the_tensor_core_path(n_coarse, n_importance, prec):
    """VERDICT r1 missing #7: --N_importance 128 (the classic NeRF 64 + 128 recipe) and the other sample counts with 192 / 256
    fine samples used to drop to the fp32 CUDA-core path (2 % of the roofline).  A 128-point tile then cuts across rays, so the
    tcgen05 kernel runs as the MLP only and k_composite (one warp per ray) does the compositing; checked like every other
    option set: coarse stage direct, fine stage teacher-forced on the oracle's z-values, 0 violations, against the oracle's
    op sequence on the GPU."""
    from nerf_sr_b200 import Renderer
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    cfg = O.RenderConfig(white_bkgd=True, N_coarse=n_coarse, N_importance=n_importance)
    pc, pf = O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17)
    rays = O.synthetic_rays(1000, 31, "blender").to(dev)                  # 1000 x 192 points: a ragged last tile
    ex = {}
    with torch.no_grad():
        ref = O.forward_rays({k: v.to(dev) for k, v in pc.items()}, {k: v.to(dev) for k, v in pf.items()}, rays, cfg, extras=ex)
    r = Renderer(cfg, dev, precision=prec)
    r.load_state_dict(0, pc)
    r.load_state_dict(1, pf)
    launches0 = r.launch_count
    out = r.forward_rays(rays, want_z_fine=True)
    assert r.launch_count - launches0 == (3 if n_coarse + n_importance > 128 else 2)      # coarse, fine MLP (+ compositing)
    well = (ex["raw_coarse"][:, -1, 3].abs() >= 1e-4) & (ex["raw_fine"][:, -1, 3].abs() >= 1e-4)
    for k in COARSE_KEYS:
        assert out[k].shape == ref[k].shape
        mx, viol = O.tolerance_violations(out[k][well].cpu(), ref[k][well].cpu())
        assert viol == 0.0, (k, mx, viol)
    assert out["fine_weights"].shape == (1000, n_coarse + n_importance) and out["z_fine"].shape == (1000, n_coarse + n_importance)
    mx, viol = O.tolerance_violations(out["z_fine"].cpu(), ex["z_fine"].cpu())           # in-kernel resampler with this n_importance
    assert float((out["z_fine"] - ex["z_fine"]).abs().median()) < 5e-6 and viol < 0.01, (mx, viol)    # z in [2, 6]: 1 ulp = 4.8e-7
    p = r.render_pass(1, rays, ex["z_fine"], want_raw=True)
    mx, viol = O.tolerance_violations(p["raw"].cpu(), ex["raw_fine"].cpu())
    assert viol == 0.0, ("raw_fine", mx, viol)
    for kl, kr in FINE_MAP:
        mx, viol = O.tolerance_violations(p[kl][well].cpu(), ref[kr][well].cpu())
        assert viol == 0.0, (kl, mx, viol)
    # end to end: the shared rule
    from conftest import e2e_bounds
    d = lambda q: {k: v.to(dev).double() for k, v in q.items()}
    with torch.no_grad():
        ref64 = O.forward_rays(d(pc), d(pf), rays.double(), cfg)
    for k in ("fine_comp_rgbs", "fine_depth"):
        _, floor = O.tolerance_violations(ref[k].cpu(), ref64[k].cpu())
        _, v64 = O.tolerance_violations(out[k].cpu(), ref64[k].cpu())
        assert v64 <= e2e_bounds(floor, 1000, prec)[0], (k, v64, floor)
    # and the fp32 CUDA-core path agrees (same option set, independent kernels)
    rs = Renderer(cfg, dev, precision="fp32_simt")
    rs.load_state_dict(0, pc)
    rs.load_state_dict(1, pf)
    ps = rs.render_pass(1, rays, ex["z_fine"])
    mx, viol = O.tolerance_violations(p["comp_rgbs"][well].cpu(), ps["comp_rgbs"][well].cpu())
    assert viol == 0.0, (mx, viol)
    rs.close()
    r.close()
