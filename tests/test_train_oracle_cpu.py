"""CPU: the training-step oracle (oracle/train_oracle.py, scope row f-1) against the committed
fixtures produced from the unmodified reference's optimize_parameters (oracle/make_golden_train.py),
against torch.optim.Adam / nn.utils.clip_grad_*, and the host-side training helpers."""
import numpy as np
import pytest
import torch

from conftest import TrainFixture, train_golden_names
from oracle import nerf_oracle as O
from oracle import ref_shim
from oracle import train_oracle as T


@pytest.mark.parametrize("name", train_golden_names())
def test_train_oracle_reproduces_golden(name):
    fx = TrainFixture(name)
    state = T.TrainState(fx.p_coarse, fx.p_fine)
    for step in range(2):
        losses, grads = T.optimize_parameters(state, fx.rays, fx.target, fx.cfg, fx.tcfg, fx.rng[step], fx.s,
                                              target_sr=fx.target_sr, ref_rays=fx.ref_rays, ref_rgbs=fx.ref_rgbs,
                                              ref_rng=fx.ref_rng[step])
        ref = fx.meta["steps"][step]
        for k, v in ref.items():
            assert abs(float(losses[k]) - v) <= 1e-5 * abs(v) + 1e-7, (step, k, float(losses[k]), v)
        for prefix, tensors in (("grad", grads), ("param", state.param_list())):
            for i, t in enumerate(tensors):
                sub, norm = fx.sub(prefix, step, i)
                flat = t.reshape(-1)
                idx = torch.from_numpy(T.golden_sub_indices(flat.numel(), i))
                # same ATen CPU kernels and op order -> bit equality in the build container; a different
                # host CPU may block its GEMMs differently, hence the hair of slack
                scale = float(sub.abs().max()) + 1e-12
                assert float((flat[idx] - sub).abs().max()) <= 2e-4 * scale, (step, prefix, i)
                assert abs(float(torch.linalg.vector_norm(flat.double())) - norm) <= 2e-4 * norm + 1e-12, (step, prefix, i)


def test_adam_step_matches_torch_optim():
    g = torch.Generator().manual_seed(3)
    ps = [torch.randn(7, 5, generator=g), torch.randn(11, generator=g)]
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    opt = torch.optim.Adam(ref, lr=5e-4, betas=(0.9, 0.999))
    m = [torch.zeros_like(p) for p in ps]
    v = [torch.zeros_like(p) for p in ps]
    for step in range(1, 5):
        grads = [torch.randn(p.shape, generator=g) * 10 ** float(torch.randn((), generator=g)) for p in ps]
        for r, gr in zip(ref, grads):
            r.grad = gr.clone()
        opt.step()
        T.adam_step(ps, grads, m, v, step, 5e-4)
        for a, b in zip(ps, ref):
            assert torch.equal(a, b.detach()), step


@pytest.mark.parametrize("kind", ["norm", "value"])
def test_clip_matches_torch(kind):
    g = torch.Generator().manual_seed(4)
    grads = [torch.randn(13, 3, generator=g), torch.randn(5, generator=g)]
    ref = [torch.nn.Parameter(torch.zeros_like(x)) for x in grads]
    for r, x in zip(ref, grads):
        r.grad = x.clone()
    tc = T.TrainConfig(grad_clip_val=0.7 if kind == "norm" else 0.3, grad_clip_type=kind)
    if kind == "norm":
        torch.nn.utils.clip_grad_norm_(ref, 0.7)
    else:
        torch.nn.utils.clip_grad_value_(ref, 0.3)
    T.clip_grads(grads, tc)
    for a, b in zip(grads, ref):
        assert torch.equal(a, b.grad)


def test_lr_schedule_matches_reference_rule():
    fx = TrainFixture(train_golden_names()[0])
    tc = T.TrainConfig(**{**fx.meta["tcfg"], "n_epochs": 3, "n_epochs_decay": 4})
    for e, lr in enumerate(fx.meta["lr_schedule_n3_d4"]):
        assert abs(T.lr_at_epoch(tc, e) - lr) <= 1e-12 * tc.lr
    d = T.TrainConfig()      # reference defaults: 20 epochs, the last 10 decaying 5e-4 -> 5e-6 (exp)
    assert T.lr_at_epoch(d, 0) == pytest.approx(d.lr) and T.lr_at_epoch(d, 9) == pytest.approx(d.lr)
    assert T.lr_at_epoch(d, 10) < d.lr and T.lr_at_epoch(d, 20) == pytest.approx(d.lr_final)


def test_sharded_gradients_average_to_the_full_batch():
    """DDP semantics used for multi-GPU training: with equal shards at LR-pixel
    granularity, the mean over ranks of the per-shard gradients IS the full-batch gradient."""
    cfg = O.RenderConfig(white_bkgd=True)
    tc = T.TrainConfig()
    pc, pf = O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17)
    rays = O.synthetic_rays(32, 9, "blender")
    target = torch.rand(8, 3, generator=torch.Generator().manual_seed(1))
    _, gc, gf, _ = T.loss_and_grads(pc, pf, rays, target, cfg, tc, None, 2)
    halves = [T.loss_and_grads(pc, pf, rays[i * 16:(i + 1) * 16], target[i * 4:(i + 1) * 4], cfg, tc, None, 2) for i in (0, 1)]
    for k in gc:
        avg = 0.5 * (halves[0][1][k] + halves[1][1][k])
        assert torch.allclose(avg, gc[k], rtol=1e-4, atol=1e-7 + 1e-4 * float(gc[k].abs().max())), k
        avg = 0.5 * (halves[0][2][k] + halves[1][2][k])
        assert torch.allclose(avg, gf[k], rtol=1e-4, atol=1e-7 + 1e-4 * float(gf[k].abs().max())), k


def test_frozen_slices_follow_the_fix_layers_regex():
    from nerf_sr_b200.training import frozen_slices
    from nerf_sr_b200.renderer import state_dict_order
    cfg = O.RenderConfig()
    names = state_dict_order(cfg.D)
    shapes = O.make_mlp_params(cfg, 1)
    numels = [shapes[n].numel() for n in names]
    assert frozen_slices(names, numels, None) == [] and frozen_slices(names, numels, "nothing") == []
    sl = frozen_slices(names, numels, r"xyz_encoding_[1-4]\.")
    assert sl == [(0, sum(numels[:8]))]                              # layers 1-4 (weight + bias each) are contiguous: merged
    sl = frozen_slices(names, numels, r"(sigma|rgb)")
    offs = [sum(numels[:i]) for i in range(len(names) + 1)]
    want = [(offs[names.index("sigma.weight")], offs[names.index("sigma.bias") + 1]),
            (offs[names.index("rgb.0.weight")], offs[names.index("rgb.0.bias") + 1])]
    # sigma.* and rgb.* are adjacent in state_dict order -> one merged range; otherwise two
    assert sl == ([(want[0][0], want[1][1])] if want[0][1] == want[1][0] else want)
    # re.match anchors at the start, like the reference: 'encoding' alone matches nothing
    assert frozen_slices(names, numels, "encoding") == []
    # the mirror of requires_grad=False: Adam on an identically zero gradient never moves a parameter
    p, m, v = [torch.randn(5)], [torch.zeros(5)], [torch.zeros(5)]
    before = p[0].clone()
    for step in (1, 2, 3):
        T.adam_step(p, [torch.zeros(5)], m, v, step, 1e-3)
    assert torch.equal(p[0], before) and not m[0].any() and not v[0].any()


def test_training_host_helpers():
    from nerf_sr_b200 import training as TR
    flat = torch.arange(10.)
    a, b = TR.unflatten_grads(flat, [(2, 3), (4,)])
    assert a.shape == (2, 3) and b.shape == (4,) and float(b[0]) == 6.0
    assert TR.image_bytes(129, 256) == 2 * 4 * 32768 and TR.image_bytes(128, 64) == 32768

    class Net(torch.nn.Module):      # parameter registration order of VanillaMLP (models/networks.py:149-180)
        def __init__(self):
            super().__init__()
            for i in range(2):
                setattr(self, f"xyz_encoding_{i+1}", torch.nn.Sequential(torch.nn.Linear(3, 3)))
            self.xyz_encoding_final = torch.nn.Linear(3, 3)
            self.dir_encoding = torch.nn.Sequential(torch.nn.Linear(3, 3))
            self.sigma = torch.nn.Linear(3, 1)
            self.rgb = torch.nn.Sequential(torch.nn.Linear(3, 3))
    net = Net()
    ps = TR.module_params_in_order(net)
    assert len(ps) == 12 and ps[0] is net.xyz_encoding_1[0].weight and ps[-1] is net.rgb[0].bias
    assert ps[8] is net.sigma.weight


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree only exists in the build container")
def test_train_oracle_bit_equals_live_reference():
    cfg = O.RenderConfig(white_bkgd=True)
    tc = T.TrainConfig()
    model, _ = ref_shim.load_reference_model("nerf_downX", ["--white_bkgd"], train=True)
    pc, pf = O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17)
    ref_shim.set_weights(model, pc, pf)
    rays = O.synthetic_rays(32, 5, "blender")
    target = torch.rand(8, 3, generator=torch.Generator().manual_seed(2))
    torch.manual_seed(11)
    model.set_input({"rays": rays.clone(), "rgbs": target.clone()})
    model.optimize_parameters()
    state = T.TrainState(pc, pf)
    rng = O.RenderRng.draw(32, cfg, torch.Generator().manual_seed(11))
    losses, _ = T.optimize_parameters(state, rays, target, cfg, tc, rng, 2)
    assert torch.equal(model.loss_tot.detach(), losses["tot"])
    ref_params = list(model.netCoarse.parameters()) + list(model.netFine.parameters())
    for a, b in zip(ref_params, state.param_list()):
        assert torch.equal(a.detach(), b)
