"""CPU: the host logic of ``patch_model`` (nerf_sr_b200/renderer.py, the one-line hook of INTEGRATION.md) with the
Renderer class replaced by a recording stand-in: near / far bookkeeping, the reference's train-mode draw order, which
option sets keep the reference path under autograd, and the view-direction column of the two model classes."""
import types

import pytest
import torch

from nerf_sr_b200 import renderer as R


@pytest.fixture
def fake_renderer(monkeypatch):
    calls = []

    class FakeRenderer:
        def __init__(self, opt, device=None, precision="bf16x3", viewdir_offset=3):
            self.n_coarse, self.n_importance = opt.N_coarse, opt.N_importance
            self.cfg = types.SimpleNamespace(no_dir=int(getattr(opt, "no_dir", False)), W=int(getattr(opt, "W", 256)))
            self._param_versions = [None, None]
            calls.append(("init", precision, viewdir_offset))

        def sync_from_modules(self, a, b):
            calls.append(("sync",))

        def forward_rays(self, rays, rng):
            calls.append(("forward_rays", None if rng is None else {k: tuple(v.shape) for k, v in rng.items()}))
            return {"coarse_comp_rgbs": torch.zeros(rays.shape[0], 3)}

    monkeypatch.setattr(R, "Renderer", FakeRenderer)
    return calls


def _model(cls_name="NeRFDownXModel", **opt):
    o = dict(N_coarse=64, N_importance=64, noise_std=1.0, no_dir=False)
    o.update(opt)
    m = type(cls_name, (), {})()
    m.opt, m.device, m.randomized = types.SimpleNamespace(**o), "cpu", False
    m.netCoarse, m.netFine = torch.nn.Linear(2, 2), torch.nn.Linear(2, 2)
    m.forward_rays = lambda rays: "reference path"
    return m


def test_eval_path_syncs_weights_and_keeps_near_far(fake_renderer):
    m = R.patch_model(_model())
    assert fake_renderer == [("init", "bf16x3", 3)]
    rays = torch.rand(5, 8)
    rays[0, 6], rays[0, 7] = 2.0, 6.0
    with torch.no_grad():
        out = m.forward_rays(rays)
    assert set(out) == {"coarse_comp_rgbs"} and fake_renderer[1:] == [("sync",), ("forward_rays", None)]
    # numpy shape-(1,) arrays like the reference's near[0].cpu().numpy() (models/nerf_downX_model.py:284)
    assert m.near.shape == (1,) and float(m.near[0]) == 2.0 and float(m.far[0]) == 6.0
    rays2 = torch.rand(3, 8)
    rays2[0, 6], rays2[0, 7] = 0.0, 1.0
    with torch.no_grad():
        m.forward_rays(rays2)
    assert float(m.near[0]) == 0.0 and float(m.far[0]) == 1.0           # refreshed per call, never stale


def test_randomized_draws_follow_the_reference_order(fake_renderer):
    m = R.patch_model(_model())
    m.randomized = True
    with torch.no_grad():
        m.forward_rays(torch.rand(7, 8))
    rng = fake_renderer[-1][1]
    assert list(rng) == ["u_coarse", "noise_coarse", "u_fine", "noise_fine"]           # models/utils.py:41, :210, :73, :210
    assert rng == {"u_coarse": (7, 64), "noise_coarse": (7, 64), "u_fine": (7, 64), "noise_fine": (7, 128)}
    m2 = R.patch_model(_model(noise_std=0.0, N_importance=0))
    m2.randomized = True
    with torch.no_grad():
        m2.forward_rays(torch.rand(4, 8))
    assert list(fake_renderer[-1][1]) == ["u_coarse"]


@pytest.mark.parametrize("precision,opt", [("fp32_simt", {}), ("fp16x3", {}), ("bf16x3", {"N_importance": 0}), ("bf16x3", {"no_dir": True}),
                                           ("bf16x3", {"W": 128})])
def test_option_sets_without_a_cuda_backward_keep_the_reference_path_under_autograd(fake_renderer, precision, opt):
    m = R.patch_model(_model(**opt), precision=precision)
    assert m.forward_rays(torch.rand(4, 8)) == "reference path"          # grad enabled, parameters require grad
    with torch.no_grad():
        assert isinstance(m.forward_rays(torch.rand(4, 8)), dict)       # inference still goes to the CUDA path
    for p in list(m.netCoarse.parameters()) + list(m.netFine.parameters()):
        p.requires_grad_(False)
    assert isinstance(m.forward_rays(torch.rand(4, 8)), dict)           # frozen nets: nothing to differentiate


def test_vanilla_model_reads_the_view_direction_from_column_8(fake_renderer):
    R.patch_model(_model("NeRFModel"))
    assert fake_renderer[-1] == ("init", "bf16x3", 8)                    # models/nerf_model.py:213: rays[:, 8:11]


def test_whole_frame_lifts_the_ray_chunk(fake_renderer):
    m = _model(ray_chunk=4096)
    R.patch_model(m)
    assert m.opt.ray_chunk >= 1 << 30                                   # chunk_batch now makes one call per frame
    m2 = _model(ray_chunk=4096)
    R.patch_model(m2, whole_frame=False)
    assert m2.opt.ray_chunk == 4096
    R.patch_model(_model())                                              # an opt without ray_chunk is left alone


def test_unsupported_option_set_keeps_the_reference_path(monkeypatch):
    """SURVEY 8b / INTEGRATION.md: nsr_create -> NSR_ERR_UNSUPPORTED leaves the model untouched, with one warning."""
    class Refusing:
        def __init__(self, *a, **k):
            raise R.NsrError(2, "D=3 is not supported")

    monkeypatch.setattr(R, "Renderer", Refusing)
    m = _model(ray_chunk=4096)
    ref = m.forward_rays
    with pytest.warns(RuntimeWarning, match="reference forward_rays stays"):
        assert R.patch_model(m) is m
    assert m.forward_rays is ref and m.opt.ray_chunk == 4096 and m._nsr_renderer is None

    class Broken:
        def __init__(self, *a, **k):
            raise R.NsrError(5, "CUDA error")

    monkeypatch.setattr(R, "Renderer", Broken)
    with pytest.raises(R.NsrError):                                      # anything else is a real failure: not swallowed
        R.patch_model(_model())


def test_near_far_are_lazy_and_assignable(fake_renderer):
    m = R.patch_model(_model())
    with pytest.raises(AttributeError):
        m.near                                                           # nothing rendered yet
    rays = torch.rand(4, 8)
    rays[0, 6], rays[0, 7] = 2.0, 6.0
    with torch.no_grad():
        m.forward_rays(rays)
    assert type(m).__name__ == "NeRFDownXModel"                          # the per-instance subclass keeps the class name
    assert m.near.dtype.name == "float32" and m.far.shape == (1,)
    m.far = 6.5                                                          # e.g. a harness replacing the array by a scalar
    assert m.far == 6.5 and float(m.near[0]) == 2.0
    with torch.no_grad():
        m.forward_rays(rays)
    assert float(m.far[0]) == 6.0                                        # a new call refreshes both


def test_reference_fallback_in_train_mode_is_chunked_by_the_original_ray_chunk(fake_renderer):
    m = _model(ray_chunk=3, N_importance=0)                              # coarse-only: no CUDA backward -> reference path
    seen = []

    def ref(rays):
        seen.append(rays.shape[0])
        return {"coarse_comp_rgbs": rays[:, :3] * 2}
    m.forward_rays = ref
    R.patch_model(m)
    assert m.opt.ray_chunk >= 1 << 30
    rays = torch.rand(8, 8)
    out = m.forward_rays(rays)                                           # grad enabled, parameters require grad
    assert seen == [3, 3, 2] and torch.equal(out["coarse_comp_rgbs"], rays[:, :3] * 2)


def test_grad_mode_looks_at_both_nets(fake_renderer, monkeypatch):
    """--fix_layers may freeze all of netCoarse: the fine net still needs the autograd path."""
    from nerf_sr_b200 import training as T
    m = R.patch_model(_model())
    for p in m.netCoarse.parameters():
        p.requires_grad_(False)
    called = []
    monkeypatch.setattr(T, "module_params_in_order", lambda net: list(net.parameters()))
    monkeypatch.setattr(T.RenderFunction, "apply", staticmethod(lambda *a: called.append(a) or tuple(range(8))))
    out = m.forward_rays(torch.rand(4, 8))
    assert called and set(out) == set(T.OUT_KEYS)
