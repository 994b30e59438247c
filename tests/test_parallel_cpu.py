"""CPU, world_size 2, gloo: the N>1 host logic (ray sharding, LR gather, gradient bucket all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf_sr_b200.parallel import allreduce_mean_, epoch_indices, rank_batches, render_sharded, shard_bounds
from oracle import nerf_oracle as O


def test_shard_bounds_respect_subpixel_groups():
    for n_lr, s, world in ((10, 2, 4), (47628, 2, 8), (40000, 4, 8), (3, 2, 8)):
        n = n_lr * s * s
        b = shard_bounds(n, world, s * s)
        assert b[0][0] == 0 and b[-1][1] == n
        for (lo, hi), (lo2, _) in zip(b, b[1:] + [(n, n)]):
            assert hi == lo2 and lo % (s * s) == 0 and hi % (s * s) == 0
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= s * s
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_render(rays, s):
    # deterministic stand-in for the CUDA renderer: per-ray value, then the reference's box average
    rgb = torch.stack([rays[:, 0], rays[:, 1] * 2, rays[:, 2] + 1], 1)
    depth = rays[:, 6] + rays[:, 7]
    return O.box_average(rgb, s), O.box_average(depth, s)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = 2
        rays = O.synthetic_rays(4 * 37, 3, "blender")          # 37 LR pixels: uneven split over 2 ranks
        (rgb, depth), bounds = render_sharded(lambda r: _fake_render(r, s), rays, s)
        ref_rgb, ref_depth = _fake_render(rays, s)
        ok = torch.equal(rgb, ref_rgb) and torch.equal(depth, ref_depth)
        # gradient bucket: mean over ranks
        g = [torch.full((5, 3), float(rank + 1)), torch.full((7,), float(10 * (rank + 1)))]
        allreduce_mean_(g)
        ok = ok and torch.allclose(g[0], torch.full((5, 3), 1.5)) and torch.allclose(g[1], torch.full((7,), 15.0))
        # training-batch sharding: every rank derives its epoch indices alone; together they are a permutation (+ wrap pad)
        mine = epoch_indices(101, world, rank, epoch=3, seed=7)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        allidx = torch.stack(parts, 1).reshape(-1)                  # interleave back: rank r holds entries r, r+world, ...
        ok = ok and allidx.numel() == 102 and sorted(allidx[:101].tolist()) == list(range(101)) and int(allidx[101]) == int(allidx[0])
        q.put((rank, bool(ok), bounds))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_render_and_grad_bucket():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == [(0, 76), (76, 148)]


@pytest.mark.parametrize("n,world,shuffle,drop_last", [(101, 2, True, False), (47628, 8, True, False), (10, 4, False, False),
                                                       (3, 8, True, False), (101, 4, True, True), (64, 8, True, True)])
def test_epoch_indices_equal_torch_distributed_sampler(n, world, shuffle, drop_last):
    """The reference's DDP loader is torch's DistributedSampler(seed=opt.seed) (data/__init__.py:94-101)."""
    from torch.utils.data import DistributedSampler
    data = list(range(n))
    for epoch in (0, 5):
        for rank in range(world):
            ref = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=shuffle, seed=11, drop_last=drop_last)
            ref.set_epoch(epoch)
            got = epoch_indices(n, world, rank, epoch, seed=11, shuffle=shuffle, drop_last=drop_last)
            assert got.tolist() == list(iter(ref)), (epoch, rank)


def test_rank_batches_follow_the_ddp_loader():
    from torch.utils.data import DataLoader, DistributedSampler
    n, batch, world = 1000, 64, 4
    data = torch.arange(n)
    for rank in range(world):
        samp = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=True, seed=0)
        samp.set_epoch(2)
        ref = [b.tolist() for b in DataLoader(data, batch_size=batch // world, sampler=samp, drop_last=True)]
        got = [b.tolist() for b in rank_batches(n, batch, world, rank, epoch=2, seed=0)]
        assert got == ref and all(len(b) == 16 for b in got)
        ref_keep = [b.tolist() for b in DataLoader(data, batch_size=batch // world, sampler=samp, drop_last=False)]
        assert [b.tolist() for b in rank_batches(n, batch, world, rank, epoch=2, seed=0, keep_last=True)] == ref_keep
    with pytest.raises(ValueError):
        list(rank_batches(n, 66, 4, 0))
    with pytest.raises(ValueError):
        epoch_indices(10, 2, 2)


# ---- patch_model under DistributedDataParallel (the reference's --accelerator ddp, models/networks.py:72-86) ----
class _TinyNet(torch.nn.Module):
    """Parameter names of the reference's VanillaMLP with D=1 (models/networks.py:149-180), tiny shapes."""

    def __init__(self):
        super().__init__()
        nn = torch.nn
        self.xyz_encoding_1 = nn.Sequential(nn.Linear(3, 4), nn.ReLU(True))
        self.xyz_encoding_final = nn.Linear(4, 4)
        self.dir_encoding = nn.Sequential(nn.Linear(4, 2), nn.ReLU(True))
        self.sigma = nn.Linear(4, 1)
        self.rgb = nn.Sequential(nn.Linear(2, 3), nn.Sigmoid())


class _RecordingRenderer:
    """Stand-in for the CUDA renderer: outputs depend on nothing, the 'backward' returns rank-dependent flat gradients."""

    def __init__(self, opt, device=None, precision="bf16x3", viewdir_offset=3):
        import types
        self.n_coarse, self.n_importance, self.cfg = opt.N_coarse, opt.N_importance, types.SimpleNamespace(no_dir=0, W=256)
        self._param_versions = [None, None]
        self.numel = None

    def load_params(self, which, params):
        self.numel = sum(p.numel() for p in params)

    def new_train_workspace(self, n):
        return torch.empty(1)

    def render_train(self, rays, rng, ws=None):
        n = rays.shape[0]
        from nerf_sr_b200.training import OUT_KEYS
        return {k: torch.zeros(n, 3 if "rgbs" in k else (4 if "weights" in k else 1)).squeeze(-1) for k in OUT_KEYS}

    def backward(self, rays, rng, grads, ws=None):
        r = dist.get_rank()
        return torch.full((self.numel,), float(r + 1)), torch.full((self.numel,), float(10 * (r + 1)))


def _ddp_worker(rank, world, port, q):
    import types
    from torch.nn.parallel import DistributedDataParallel as DDP
    from nerf_sr_b200 import renderer as R
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        R.Renderer = _RecordingRenderer
        torch.manual_seed(0)
        m = type("NeRFDownXModel", (), {})()
        m.opt = types.SimpleNamespace(N_coarse=64, N_importance=64, noise_std=0.0, ray_chunk=4096)
        m.device, m.randomized = "cpu", False
        m.netCoarse, m.netFine = DDP(_TinyNet()), DDP(_TinyNet())
        m.forward_rays = lambda rays: None
        R.patch_model(m)
        out = m.forward_rays(torch.rand(6, 8))
        (out["coarse_comp_rgbs"].sum() + out["fine_comp_rgbs"].sum()).backward()
        gc = torch.cat([p.grad.reshape(-1) for p in m.netCoarse.parameters()])
        gf = torch.cat([p.grad.reshape(-1) for p in m.netFine.parameters()])
        q.put((rank, float(gc.min()), float(gc.max()), float(gf.min()), float(gf.max())))
    finally:
        dist.destroy_process_group()


def test_patch_model_averages_gradients_over_the_ddp_group():
    """RenderFunction bypasses DDP.forward, so its backward must do the reducer's all-reduce: both ranks
    end with the MEAN of the rank-local gradients (1.5 / 15), not their own (1 / 10 and 2 / 20)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, 1.5, 1.5, 15.0, 15.0), (1, 1.5, 1.5, 15.0, 15.0)], res
