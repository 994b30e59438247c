"""CPU: checkpoint files and the learning-rate rule (nerf_sr_b200/checkpoints.py) against torch's own schedulers, the
schedule the reference produced (stored in the training fixtures) and -- in the build container -- the reference's
save_networks / load_networks in both directions."""
import json
import os

import pytest
import torch

from conftest import TrainFixture
from nerf_sr_b200 import checkpoints as K
from oracle import nerf_oracle as O
from oracle import ref_shim


def test_lr_rule_matches_the_reference_schedule_in_the_fixtures():
    for name in ("train_step_blender", "train_step_llff_clip"):
        fx = TrainFixture(name)
        want = fx.meta["lr_schedule_n3_d4"]                     # LambdaLR driven by the reference's get_scheduler, 8 epochs
        t = fx.tcfg
        got = [K.lr_at_epoch(e, t.lr, t.lr_final, 3, 4, t.lr_policy) for e in range(len(want))]
        assert got == pytest.approx(want, rel=1e-12)
        assert all(a >= b for a, b in zip(got, got[1:]))      # t is not clamped at 1: past n_epochs the rule keeps extrapolating
        assert K.lr_at_epoch(9, t.lr, t.lr_final, 20, 10, t.lr_policy) == pytest.approx(t.lr, rel=1e-12)   # held for n_epochs - n_epochs_decay


@pytest.mark.parametrize("policy", ["linear", "exp", "step"])
def test_lr_schedule_object_matches_torch_schedulers(policy):
    lr, lr_final, n_epochs, n_decay = 5e-4, 5e-6, 6, 4
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([{"params": [p], "initial_lr": lr}], lr=lr)
    for start in (0, 3):                                          # fresh run and --continue_train from epoch 3
        for g in opt.param_groups:
            g["lr"] = lr
        if policy == "step":
            sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.1, last_epoch=start - 1)
        else:
            def rule(epoch):
                t = max(0, epoch + 1 - n_epochs + n_decay) / float(n_decay + 1)
                cur = lr * (1 - t) + lr_final * t if policy == "linear" else \
                    torch.exp(torch.tensor(torch.log(torch.tensor(lr, dtype=torch.float64)) * (1 - t)
                                           + torch.log(torch.tensor(lr_final, dtype=torch.float64)) * t)).item()
                return cur / lr
            sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=rule, last_epoch=start - 1)
        mine = K.LrSchedule(lr, lr_final, n_epochs, n_decay, policy, lr_decay_epochs=2, lr_decay_gamma=0.1, current_epoch=start)
        for _ in range(8):
            assert mine.lr == pytest.approx(opt.param_groups[0]["lr"], rel=1e-12), (policy, start, mine.epoch)
            opt.step()
            sched.step()
            mine.step()
    with pytest.raises(NotImplementedError):
        K.LrSchedule(lr_policy="cosine")


def test_checkpoint_roundtrip_and_filters(tmp_path):
    cfg = O.RenderConfig()
    pc, pf = O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17)
    d = str(tmp_path)
    a, b = K.save_networks(d, 7, pc, pf)
    assert os.path.basename(a) == "7_net_Coarse.pth" and os.path.basename(b) == "7_net_Fine.pth"
    K.save_networks(d, 12, pf, pc)
    K.save_networks(d, "latest", pc, pf)
    assert K.latest_epoch(d) == 12                                 # 'latest' files are ignored, like the reference's glob
    lc, lf = K.load_networks(d, 7)
    assert list(lc) == list(pc) and all(torch.equal(lc[k], pc[k]) for k in pc) and all(torch.equal(lf[k], pf[k]) for k in pf)
    sub, _ = K.load_networks(d, 7, keys=r"xyz_encoding_[12]\.")
    assert sorted(sub) == ["xyz_encoding_1.0.bias", "xyz_encoding_1.0.weight", "xyz_encoding_2.0.bias", "xyz_encoding_2.0.weight"]
    torch.save({"module." + k: v for k, v in pc.items()}, os.path.join(d, "3_net_Coarse.pth"))
    torch.save(dict(pf), os.path.join(d, "3_net_Fine.pth"))
    lc3, _ = K.load_networks(d, 3)
    assert list(lc3) == list(pc)
    with pytest.raises(FileNotFoundError):
        K.load_networks(d, 99)
    with pytest.raises(FileNotFoundError):
        K.latest_epoch(os.path.join(d, "nowhere"))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree only exists in the build container")
def test_checkpoints_interchange_with_the_reference(tmp_path):
    cfg = O.RenderConfig(white_bkgd=True)
    pc, pf = O.make_mlp_params(cfg, 4), O.make_mlp_params(cfg, 17)
    model, opt = ref_shim.load_reference_model("nerf_downX", ["--white_bkgd"], train=True)
    ref_shim.set_weights(model, pc, pf)
    model.save_dir = os.path.join(str(tmp_path), "ref_exp")
    os.makedirs(model.save_dir)
    model.save_networks(5)                                          # reference writes -> we read
    lc, lf = K.load_networks(model.save_dir, 5)
    assert list(lc) == list(pc) and all(torch.equal(lc[k], pc[k]) for k in pc) and all(torch.equal(lf[k], pf[k]) for k in pf)
    assert K.latest_epoch(model.save_dir) == 5
    pc2, pf2 = O.make_mlp_params(cfg, 31), O.make_mlp_params(cfg, 34)
    opt.checkpoints_dir = str(tmp_path)                             # we write -> reference reads
    K.save_networks(os.path.join(str(tmp_path), "mine_exp"), 9, pc2, pf2)
    model.load_networks("mine_exp", 9)
    unwrap = lambda n: n.module if hasattr(n, "module") else n
    for net, want in ((unwrap(model.netCoarse), pc2), (unwrap(model.netFine), pf2)):
        sd = net.state_dict()
        assert list(sd) == list(want) and all(torch.equal(sd[k].cpu(), want[k]) for k in want)
