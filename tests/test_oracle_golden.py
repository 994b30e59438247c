"""CPU: the oracle restatement against the committed golden vectors (which were
produced by the unmodified reference, oracle/make_golden.py) and -- when the
reference tree is present (build container only) -- against the reference live."""
import numpy as np
import pytest
import torch

from conftest import golden_names
from oracle import nerf_oracle as O
from oracle import ref_shim


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_golden(name, load_fixture):
    fx = load_fixture(name)
    with torch.no_grad():
        extras = {}
        out = O.forward_rays(fx.p_coarse, fx.p_fine, fx.rays, fx.cfg, fx.rng, extras=extras)
    assert set(out) == set(fx.out)
    for k, ref in fx.out.items():
        # same ATen CPU kernels, same op order, same image -> expect bit equality;
        # allow a hair for a different host CPU's GEMM blocking.
        max_abs, viol = O.tolerance_violations(out[k], ref, rtol=1e-5, atol=1e-6)
        floor = fx.meta["fp64_floor"][k]
        assert viol <= max(0.0, floor["viol"]) + 0.02, (k, max_abs, viol)
    assert torch.allclose(extras["z_coarse"], fx.z_coarse, rtol=0, atol=0)


@pytest.mark.parametrize("name", ["eval_blender", "eval_llff"])
def test_teacher_forced_fine_matches(name, load_fixture):
    """Protocol (ii): feeding the golden fine z-values must reproduce fine_* tightly."""
    fx = load_fixture(name)
    with torch.no_grad():
        out = O.forward_rays(fx.p_coarse, fx.p_fine, fx.rays, fx.cfg, None, z_fine_override=fx.z_fine)
    for k in ("fine_comp_rgbs", "fine_depth", "fine_opacity", "fine_weights"):
        max_abs, viol = O.tolerance_violations(out[k], fx.out[k], rtol=1e-5, atol=1e-6)
        assert viol == 0.0, (k, max_abs)


def test_fixtures_are_not_degenerate(load_fixture):
    for name in golden_names():
        fx = load_fixture(name)
        if fx.cfg.sigma_activation == "softplus":
            continue   # softplus sigma>0 and delta_last=1e10 force opacity==1
        for k, v in fx.meta["mean_opacity"].items():
            assert 0.02 < v < 0.999, (name, k, v)


def test_param_count_matches_survey():
    shapes = O.mlp_param_shapes(O.RenderConfig())
    n = sum(int(np.prod(s)) for _, s in shapes)
    assert n == 595844          # SURVEY.md section 8a layer table
    macs = sum(int(np.prod(s)) for k, s in shapes if k.endswith("weight"))
    assert macs == 593408


def test_posenc_layout():
    x = torch.tensor([[0.25, -1.5, 3.0]])
    e = O.posenc(x, 10)
    assert e.shape == (1, 63)
    assert torch.equal(e[:, :3], x)
    assert torch.allclose(e[:, 3:6], torch.sin(x)) and torch.allclose(e[:, 6:9], torch.cos(x))
    assert torch.allclose(e[:, 9:12], torch.sin(2 * x))
    assert O.posenc(x, 4).shape == (1, 27)


def test_box_average_groups_contiguous_subpixels():
    x = torch.arange(32, dtype=torch.float32).view(16, 2)
    y = O.box_average(x, 2)
    assert y.shape == (4, 2)
    assert torch.equal(y[0], x[:4].mean(0))


def test_raygen_golden():
    z = np.load("tests/golden/raygen.npz") if False else np.load(
        __import__("os").path.join(__import__("conftest").GOLDEN_DIR, "raygen.npz"))
    for tag in ("blender", "llff"):
        H, W, s, focal, ndc, near, far = z[f"{tag}_params"]
        rays = O.build_frame_rays(torch.from_numpy(z[f"{tag}_c2w"]), int(H), int(W), float(focal),
                                  int(s), float(near), float(far), bool(ndc))
        assert torch.equal(rays, torch.from_numpy(z[f"{tag}_rays"]))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree only exists in the build container")
def test_oracle_bit_equals_live_reference(load_fixture):
    fx = load_fixture("eval_llff")
    model, _ = ref_shim.load_reference_model("nerf_downX", fx.meta["reference_args"])
    ref_shim.set_weights(model, fx.p_coarse, fx.p_fine)
    with torch.no_grad():
        ref = model.forward_rays(fx.rays[:64])
        ora = O.forward_rays(fx.p_coarse, fx.p_fine, fx.rays[:64], fx.cfg)
    for k in ref:
        assert torch.equal(ref[k], ora[k]), k
