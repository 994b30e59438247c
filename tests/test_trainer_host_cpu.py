"""CPU: the host-side orchestration of ``Trainer`` (nerf_sr_b200/training.py) against a recording stand-in for the
renderer: which C-ABI seams one iteration calls, in which order and with which arguments -- for the plain LR loss, the
fused loss epilogue (variance / depth variance / SISR), the reference-view batch, norm / value clipping, frozen layers
and the learning-rate override.  (The arithmetic behind every seam is tested on the GPU in tests/test_gpu_train.py.)"""
import types

import pytest
import torch

from nerf_sr_b200 import training as TR
from nerf_sr_b200.renderer import state_dict_order
from oracle import nerf_oracle as O

CFG = O.RenderConfig(noise_std=1.0)
PC, PF = O.make_mlp_params(CFG, 1), O.make_mlp_params(CFG, 2)
NAMES = state_dict_order(CFG.D)
NUMELS = [PC[n].numel() for n in NAMES]
TOTAL = sum(NUMELS)


class FakeLib:
    def nsr_grad_numel(self, h):
        return TOTAL

    def nsr_param_numel(self, h, i):
        return NUMELS[i]


class FakeRenderer:
    """Records every seam call; returns CPU tensors of the shapes the CUDA library produces."""

    def __init__(self):
        self.device = torch.device("cpu")
        self.cfg = types.SimpleNamespace(D=CFG.D, n_coarse=64, n_importance=64, noise_std=1.0)
        self.lib, self._h, self.calls = FakeLib(), None, []

    def load_params(self, which, params):
        self.calls.append(("load_params", which))

    def new_train_workspace(self, n):
        self.calls.append(("new_train_workspace", n))
        return torch.zeros(1)

    def render_train(self, rays, rng, want_weights=True, want_z_fine=False, ws=None):
        self.calls.append(("render_train", rays.shape[0], ws is not None))
        n = rays.shape[0]
        return {k: torch.rand(n, 3) if "rgbs" in k else torch.rand(n) for k in
                ("coarse_comp_rgbs", "coarse_depth", "coarse_opacity", "fine_comp_rgbs", "fine_depth", "fine_opacity")}

    def lr_loss_grad(self, hr, tgt, s, lam, want_grad=True, metrics_out=None):
        self.calls.append(("lr_loss_grad", s, lam))
        m = torch.tensor([lam * 0.5, 3.0])
        if metrics_out is not None:
            metrics_out.copy_(m)
        return torch.zeros(tgt.shape[0], 3), m, torch.ones_like(hr)

    def loss_epilogue(self, hr, tgt, s, lam, hr_depth=None, lambda_var=0.0, lambda_depth_var=0.0, far=0.0, target_hr=None,
                      want_grad=True, lambda_hr=1.0):
        self.calls.append(("loss_epilogue", tgt is not None, hr_depth is not None, lambda_var, lambda_depth_var, far,
                           target_hr is not None, lambda_hr))
        out = {"lr_rgb": torch.zeros(hr.shape[0] // (s * s), 3), "metrics": torch.arange(8, dtype=torch.float32), "g_rgb": torch.ones_like(hr)}
        if hr_depth is not None:
            out["g_depth"] = torch.ones(hr.shape[0])
        return out

    def backward(self, rays, rng, grads, ws=None, out=None):
        self.calls.append(("backward", rays.shape[0], tuple(sorted(grads)), ws is not None))
        if out is not None:
            out[0].fill_(1.0)
            out[1].fill_(2.0)
            return out
        return torch.ones(TOTAL), torch.full((TOTAL,), 2.0)

    def clip_coef(self, a, b, max_norm):
        self.calls.append(("clip_coef", max_norm))
        return torch.tensor([0.5, 1.0])

    def adam_step(self, params, grad, m, v, step, lr, beta1, beta2, eps, coef, clip_value):
        self.calls.append(("adam_step", step, lr, coef is not None, clip_value, float(grad.sum())))


def _names(calls):
    return [c[0] for c in calls]


def test_plain_iteration_call_sequence():
    r = FakeRenderer()
    tr = TR.Trainer(r, PC, PF, lr=1e-3, lambda_coarse_mse=0.5, downscale=2)
    assert _names(r.calls) == ["load_params", "load_params"]
    r.calls.clear()
    rays, tgt = torch.rand(64, 8), torch.rand(16, 3)
    m = tr.optimize_parameters(rays, tgt)
    assert _names(r.calls) == ["render_train", "lr_loss_grad", "lr_loss_grad", "backward", "adam_step", "load_params",
                               "adam_step", "load_params"]
    assert r.calls[1] == ("lr_loss_grad", 2, 0.5) and r.calls[2] == ("lr_loss_grad", 2, 1.0)
    assert r.calls[3] == ("backward", 64, ("coarse_comp_rgbs", "fine_comp_rgbs"), False)
    assert r.calls[4][:5] == ("adam_step", 1, 1e-3, False, 0.0) and r.calls[6][1] == 1      # one step count for both nets
    assert m.shape == (4,) and tr.last_terms is None
    r.calls.clear()
    tr.optimize_parameters(rays, tgt, lr=2e-4)                          # per-epoch schedule: the caller passes the rate
    assert r.calls[4][1:3] == (2, 2e-4)


def test_fused_epilogue_is_used_when_any_extra_term_is_on():
    r = FakeRenderer()
    tr = TR.Trainer(r, PC, PF, downscale=2, lambda_coarse_var=0.02, lambda_fine_depth_var=0.05)
    r.calls.clear()
    rays = torch.rand(64, 8)
    rays[0, 7] = 6.0
    tr.optimize_parameters(rays, torch.rand(16, 3), target_sr=torch.rand(64, 3))
    assert _names(r.calls)[:4] == ["render_train", "loss_epilogue", "loss_epilogue", "backward"]
    # (has LR target, has depth, lambda_var, lambda_depth_var, far, has HR target, lambda_hr); far defaults to rays[0, 7]
    assert r.calls[1][1:] == (True, False, 0.02, 0.0, 6.0, True, 1.0)
    assert r.calls[2][1:] == (True, True, 0.0, 0.05, 6.0, True, 1.0)
    assert r.calls[3][2] == ("coarse_comp_rgbs", "fine_comp_rgbs", "fine_depth")      # depth gradient only where its term is on
    assert tr.last_terms.shape == (2, 8) and tr.last_metrics.tolist() == [0.0, 1.0, 0.0, 1.0]
    r.calls.clear()
    tr.optimize_parameters(rays, torch.rand(16, 3), far=1.0)           # an explicit far plane avoids the host read
    assert r.calls[2][5] == 1.0


def test_reference_view_batch_runs_a_second_forward_and_sums_the_gradients():
    r = FakeRenderer()
    tr = TR.Trainer(r, PC, PF, downscale=2)
    r.calls.clear()
    tr.optimize_parameters(torch.rand(64, 8), torch.rand(16, 3), ref_rays=torch.rand(40, 8), ref_rgbs=torch.rand(40, 3))
    assert _names(r.calls)[:10] == ["new_train_workspace", "new_train_workspace", "render_train", "render_train", "loss_epilogue",
                                    "loss_epilogue", "loss_epilogue", "loss_epilogue", "backward", "backward"]
    assert r.calls[2] == ("render_train", 64, True) and r.calls[3] == ("render_train", 40, True)    # private stash each
    assert r.calls[6][1:] == (False, False, 0.0, 0.0, 0.0, True, 0.25)                              # no LR target, MSE / s^2
    assert r.calls[8][1] == 64 and r.calls[9][1] == 40 and r.calls[8][3] and r.calls[9][3]
    adam = [c for c in r.calls if c[0] == "adam_step"]
    assert adam[0][5] == 2.0 * TOTAL and adam[1][5] == 4.0 * TOTAL                                # main + reference gradients
    assert tr.last_ref_terms.shape == (2,)
    with pytest.raises(TR.NsrError):
        tr.optimize_parameters(torch.rand(64, 8), torch.rand(16, 3), ref_rays=torch.rand(40, 8), ref_rgbs=torch.rand(39, 3))


@pytest.mark.parametrize("kind", ["norm", "value"])
def test_clipping_and_frozen_layers(kind):
    r = FakeRenderer()
    tr = TR.Trainer(r, PC, PF, downscale=2, grad_clip_val=0.1, grad_clip_type=kind, fix_layers=r"xyz_encoding_[1-4]\.")
    r.calls.clear()
    tr.optimize_parameters(torch.rand(64, 8), torch.rand(16, 3))
    frozen = sum(NUMELS[:8])
    adam = [c for c in r.calls if c[0] == "adam_step"]
    assert adam[0][5] == float(TOTAL - frozen) and adam[1][5] == 2.0 * (TOTAL - frozen)            # frozen slices zeroed first
    if kind == "norm":
        assert "clip_coef" in _names(r.calls) and adam[0][3] is True and adam[0][4] == 0.0
        assert _names(r.calls).index("clip_coef") < _names(r.calls).index("adam_step")
    else:
        assert "clip_coef" not in _names(r.calls) and adam[0][3] is False and adam[0][4] == 0.1


def test_draw_rng_follows_the_reference_order_and_shapes():
    r = FakeRenderer()
    tr = TR.Trainer(r, PC, PF, downscale=2)
    g = torch.Generator().manual_seed(3)
    rng = tr.draw_rng(10, g)
    assert list(rng) == ["u_coarse", "noise_coarse", "u_fine", "noise_fine"]
    assert [tuple(v.shape) for v in rng.values()] == [(10, 64), (10, 64), (10, 64), (10, 128)]
    g2 = torch.Generator().manual_seed(3)                               # same generator stream as the oracle's draw
    ref = O.RenderRng.draw(10, CFG, g2)
    assert torch.equal(rng["u_coarse"], ref.u_coarse) and torch.equal(rng["noise_fine"], ref.noise_fine)
    r.cfg.noise_std = 0.0
    assert list(tr.draw_rng(4, g)) == ["u_coarse", "u_fine"]


def test_trainer_kwargs_from_a_reference_style_opt():
    opt = types.SimpleNamespace(lr=1e-3, beta1=0.8, lambda_coarse_mse=0.5, lambda_fine_mse=1.0, grad_clip_val=0.05, grad_clip_type="value",
                                downscale=4, use_var_loss=True, lambda_coarse_var=0.02, lambda_fine_var=0.03, use_depth_var_loss=False,
                                lambda_coarse_depth_var=0.07, lambda_fine_depth_var=0.07, fix_layers=r"dir_")
    kw = TR.trainer_kwargs_from_opt(opt)
    assert kw == dict(lr=1e-3, beta1=0.8, lambda_coarse_mse=0.5, lambda_fine_mse=1.0, grad_clip_val=0.05, grad_clip_type="value", downscale=4,
                      lambda_coarse_var=0.02, lambda_fine_var=0.03, lambda_coarse_depth_var=0.0, lambda_fine_depth_var=0.0, fix_layers=r"dir_")
    r = FakeRenderer()
    tr = TR.Trainer(r, PC, PF, **kw)                                  # every key is a Trainer argument
    assert tr.lam == (0.5, 1.0) and tr.lam_var == (0.02, 0.03) and tr.lam_dvar == (0.0, 0.0) and tr.s == 4 and tr.clip_type == "value"
    assert tr._frozen and tr.beta1 == 0.8
    # defaults of a bare namespace = the reference's defaults with every optional term off
    kw0 = TR.trainer_kwargs_from_opt(types.SimpleNamespace())
    assert kw0["lr"] == 5e-4 and kw0["lambda_coarse_var"] == 0.0 and kw0["grad_clip_val"] == 0.0 and kw0["fix_layers"] is None


def test_trainer_kwargs_from_the_live_reference_options():
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree only exists in the build container")
    _, opt = ref_shim.load_reference_model("nerf_downX", ["--use_var_loss", "--lambda_fine_var", "0.5", "--grad_clip_val", "0.1",
                                                          "--downscale", "2", "--lr", "2e-4"], train=True)
    kw = TR.trainer_kwargs_from_opt(opt)
    assert kw["lr"] == 2e-4 and kw["beta1"] == opt.beta1 and kw["downscale"] == 2 and kw["grad_clip_val"] == 0.1
    assert kw["lambda_coarse_var"] == opt.lambda_coarse_var == 0.01 and kw["lambda_fine_var"] == 0.5
    assert kw["lambda_coarse_depth_var"] == 0.0 and kw["lambda_coarse_mse"] == opt.lambda_coarse_mse
