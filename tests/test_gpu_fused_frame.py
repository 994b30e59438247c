"""GPU: the one-launch frame kernel (k_tc_pass<.., FUSED>: coarse pass + resampling + fine pass, optionally ray generation
in the front-end and the s x s box average in the compositing epilogue; VERDICT r1 item 8 / north_star's "all fused")
against the separate launches it replaces (coarse k_tc_pass, fine k_tc_pass, k_box_average, k_generate_rays).

Both paths run the same device functions in the same order per ray, so the comparison is BIT-EXACT (torch.equal) on every
output -- composites, depths, opacities, per-sample weights, merged z-values, LR images.  Parity of either path against the
oracle / the reference is the rest of the GPU suite's job (it runs on the one-launch path wherever the option set allows).
`set_debug_flags(64)` keeps the separate launches on the same handle."""
import pytest
import torch

from nerf_sr_b200 import synthetic as S

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _renderer(prec, white=True, noise_std=0.0):
    from nerf_sr_b200 import Renderer
    cfg = S.RenderConfig(white_bkgd=white, noise_std=noise_std)
    r = Renderer(cfg, torch.device(DEV), precision=prec)
    r.load_state_dict(0, S.make_mlp_params(cfg, 4, sigma_bias=0.3, bias_std=0.05))
    r.load_state_dict(1, S.make_mlp_params(cfg, 17, sigma_bias=0.3, bias_std=0.05))
    return r


def _both(r, fn):
    """fn() under the one-launch path and under the separate launches: (out_fused, launches_fused, out_sep, launches_sep)."""
    r.set_debug_flags(0)
    l0 = r.launch_count
    a = fn()
    l1 = r.launch_count
    r.set_debug_flags(64)
    b = fn()
    l2 = r.launch_count
    r.set_debug_flags(0)
    torch.cuda.synchronize()
    return a, l1 - l0, b, l2 - l1


def _assert_identical(a, b):
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert torch.isfinite(a[k]).all(), k
        assert torch.equal(a[k], b[k]), (k, float((a[k] - b[k]).abs().max()))


# 1 and 2 rays: a single ray pair per CTA (the fine tile depends on the tile right before it); 5 / 131: odd counts (a
# half-filled coarse tile); 300: one unit per CTA, CTA pairs with a dummy unit; 1500: several units per CTA, skewed order
@pytest.mark.parametrize("n", [1, 2, 5, 131, 300, 1500])
@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3"])
def test_forward_rays_in_one_launch_is_bit_identical(prec, n):
    r = _renderer(prec)
    rays = S.synthetic_rays(n, 31 + n, "blender").to(DEV)
    a, la, b, lb = _both(r, lambda: r.forward_rays(rays, want_weights=True, want_z_fine=True))
    assert (la, lb) == (1, 2)
    _assert_identical(a, b)
    r.close()


@pytest.mark.parametrize("kind,white", [("blender", True), ("llff", False)])
def test_train_mode_draws_in_one_launch_are_bit_identical(kind, white):
    n = 777
    r = _renderer("bf16x3", white=white, noise_std=1.0)
    rays = S.synthetic_rays(n, 5, kind).to(DEV)
    g = torch.Generator().manual_seed(9)
    rng = {"u_coarse": torch.rand(n, 64, generator=g), "noise_coarse": torch.randn(n, 64, generator=g),
           "u_fine": torch.rand(n, 64, generator=g), "noise_fine": torch.randn(n, 128, generator=g)}
    a, la, b, lb = _both(r, lambda: r.forward_rays(rays, rng=rng, want_weights=True, want_z_fine=True))
    assert (la, lb) == (1, 2)
    _assert_identical(a, b)
    r.close()


# The box average runs in the kernel's epilogue when that costs the busiest CTA < 3 % more tiles than ray-pair granularity
# would (nsr_api.cu: lr_in_kernel_pays) -- always for frames, and here for the batch sizes that divide evenly over the 148 SMs
# (1184 LR pixels at s = 2, 296 at s = 4) -- or when the caller does not want the HR composite at all; otherwise the frame
# kernel is followed by k_box_average launches.  Either way the results are the same bits.
@pytest.mark.parametrize("s,n_lr,in_kernel", [(2, 1, False), (2, 75, True), (2, 1184, True), (2, 1201, False), (4, 1, False),
                                              (4, 37, False), (4, 296, True)])
def test_box_average_in_the_compositing_epilogue_is_bit_identical(s, n_lr, in_kernel):
    r = _renderer("bf16x3")
    n = n_lr * s * s
    rays = S.synthetic_rays(n, 100 + n_lr, "blender").to(DEV)
    a, la, b, lb = _both(r, lambda: r.render_frame(rays, s))
    assert lb == 2 + 4                             # coarse, fine, 4 x k_box_average
    pays = bool(r.lib.nsr_debug_frame_lr_in_kernel(r._h, n, s))          # the library's own rule for this batch on this GPU
    assert la == (1 if pays else 1 + 4)
    if torch.cuda.get_device_properties(0).multi_processor_count == 148:
        assert pays == in_kernel
    _assert_identical(a, b)
    for name in ("coarse", "fine"):                # and the LR image is the public box average of the HR image
        assert torch.equal(a[f"{name}_lr_rgb"], r.box_average(a[f"{name}_comp_rgbs"], s))
        assert torch.equal(a[f"{name}_lr_depth"], r.box_average(a[f"{name}_depth"], s).reshape(-1))
    l0 = r.launch_count
    lr_only = r.render_frame(rays, s, want_hr=False)             # HR composites not written at all: always in the kernel
    assert r.launch_count - l0 == 1
    assert set(lr_only) == {"coarse_lr_rgb", "coarse_lr_depth", "fine_lr_rgb", "fine_lr_depth"}
    for k, v in lr_only.items():
        assert torch.equal(v, a[k]), k
    r.close()


@pytest.mark.parametrize("H,W,s,ndc", [(24, 32, 2, False), (16, 16, 4, False), (36, 28, 2, True), (20, 12, 1, False)])
def test_rays_generated_in_the_front_end_are_bit_identical(H, W, s, ndc):
    r = _renderer("bf16x3", white=not ndc)
    c2w = torch.tensor([[0.96, -0.10, 0.26, 1.1], [0.05, 0.98, 0.19, 0.7], [-0.27, -0.17, 0.95, 3.6]])
    kw = dict(pose=c2w, H=H, W=W, focal=0.9 * W, ndc=ndc, near=2.0, far=6.0)
    a, la, b, lb = _both(r, lambda: r.render_frame(None, s, **kw))
    assert lb == 1 + 2 + (4 if s > 1 else 0)                     # k_generate_rays, coarse, fine, box averages
    assert la in (1, 1 + 4) and la < lb                          # small rasters: box averages after the frame kernel
    _assert_identical(a, b)
    # ... and equal to the public pieces called one by one
    rays = r.generate_rays(c2w, H, W, 0.9 * W, s=s, ndc=ndc, near=2.0, far=6.0)
    r.set_debug_flags(64)
    ref = r.forward_rays(rays, want_weights=False)
    r.set_debug_flags(0)
    for k in ("coarse_comp_rgbs", "coarse_depth", "fine_comp_rgbs", "fine_depth", "fine_opacity"):
        assert torch.equal(a[k], ref[k]), k
    r.close()


def test_host_frame_pipeline_in_one_launch_per_chunk_is_bit_identical():
    r = _renderer("bf16x3")
    sm148 = torch.cuda.get_device_properties(0).multi_processor_count == 148
    for n_lr, la_want in ((1184, 1), (900, 3)):                  # in the epilogue | frame kernel + 2 x k_box_average
        rays = S.synthetic_rays(4 * n_lr, 3, "blender")
        a, la, b, lb = _both(r, lambda: dict(zip(("rgb", "depth"), r.render_frame_host(rays, 2))))
        assert lb == 4 and (la == la_want or not sm148)
        _assert_identical(a, b)
    c2w = torch.tensor([[1.0, 0.0, 0.0, 0.2], [0.0, 1.0, 0.0, -0.1], [0.0, 0.0, 1.0, 4.0]])
    a, la, b, lb = _both(r, lambda: dict(zip(("rgb", "depth"), r.render_pose_host(c2w, 64, 74, 50.0, s=2))))
    assert lb == 5 and (la == 1 or not sm148)                    # 64 x 74 = 4 x 1184 rays, no ray buffer, no HR composite
    _assert_identical(a, b)
    r.close()


def test_pose_frames_larger_than_one_host_chunk_are_bit_identical():
    """A frame of more than 262 144 rays goes through the host pipeline in several chunks; with in-kernel ray generation
    every chunk passes its first row (`rg_first`) to the kernel.  600 x 600 at s = 2 = 360 000 rays = two chunks."""
    r = _renderer("bf16x3")
    c2w = torch.tensor([[0.96, -0.10, 0.26, 1.1], [0.05, 0.98, 0.19, 0.7], [-0.27, -0.17, 0.95, 3.6]])
    a, la, b, lb = _both(r, lambda: dict(zip(("rgb", "depth"), r.render_pose_host(c2w, 600, 600, 540.0, s=2))))
    assert la == 2 and lb == 1 + 2 * 4                           # one launch per chunk | rays + (coarse, fine, 2 box averages) per chunk
    _assert_identical(a, b)
    # the device-side entry point on the same pose: one launch for the whole frame, the same LR image
    o = r.render_frame(None, 2, pose=c2w, H=600, W=600, focal=540.0, want_hr=False)
    assert torch.equal(o["fine_lr_rgb"].cpu(), a["rgb"]) and torch.equal(o["fine_lr_depth"].cpu(), a["depth"])
    r.close()


def test_single_cta_launches_of_the_frame_kernel_are_bit_identical_too():
    """NSR_TC_CLUSTER=1 (the A/B switch that launches single CTAs instead of CTA pairs: CTAs then run different numbers of
    units and nobody multicasts) is read at nsr_create, so the check runs in a child process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, NSR_TC_CLUSTER="1")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", os.path.abspath(__file__), "-k",
           "(forward_rays_in_one_launch and bf16x3) or (box_average and 1184) or larger_than_one_host_chunk"]
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    assert " passed" in p.stdout and "failed" not in p.stdout, p.stdout[-1000:]


def test_option_sets_outside_the_one_launch_kernel_still_render_frames():
    """nsr_render_frame is a complete entry point: 64 + 128 samples (MLP-only fine pass) and the fp32 path fall back to the
    separate launches inside the library and give the same answers as the public pieces."""
    from nerf_sr_b200 import Renderer
    for prec, nimp in (("bf16x3", 128), ("fp32_simt", 64)):
        cfg = S.RenderConfig(white_bkgd=True, N_importance=nimp)
        r = Renderer(cfg, torch.device(DEV), precision=prec)
        r.load_state_dict(0, S.make_mlp_params(cfg, 4))
        r.load_state_dict(1, S.make_mlp_params(cfg, 17))
        c2w = torch.tensor([[1.0, 0.0, 0.0, 0.2], [0.0, 1.0, 0.0, -0.1], [0.0, 0.0, 1.0, 4.0]])
        out = r.render_frame(None, 2, pose=c2w, H=16, W=24, focal=30.0)
        rays = r.generate_rays(c2w, 16, 24, 30.0, s=2)
        ref = r.forward_rays(rays, want_weights=False)
        for k in ("coarse_comp_rgbs", "fine_comp_rgbs", "fine_depth"):
            assert torch.equal(out[k], ref[k]), (prec, k)
        assert torch.equal(out["fine_lr_rgb"], r.box_average(ref["fine_comp_rgbs"], 2)), prec
        lr_only = r.render_frame(rays, 2, want_hr=False)            # HR composites not requested: the library lends workspace
        assert set(lr_only) == {"coarse_lr_rgb", "coarse_lr_depth", "fine_lr_rgb", "fine_lr_depth"}
        assert torch.equal(lr_only["fine_lr_rgb"], out["fine_lr_rgb"]) and torch.equal(lr_only["coarse_lr_depth"], out["coarse_lr_depth"]), prec
        r.close()
