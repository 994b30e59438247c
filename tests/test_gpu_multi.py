"""GPU, >= 2 devices (skipped on a 1-GPU box): the N > 1 paths on real hardware, one process per GPU over NCCL.

  * P2PGradReducer: the CUDA-IPC wiring of libnsr_b200's one-kernel all-reduce across PROCESSES (the kernel itself is
    tested on one device in tests/test_gpu_comm.py) -- result == mean in fixed rank order, identical on every rank, and
    equal to what the NCCL collective produces up to fp32 summation order.
  * Trainer under data parallelism at the reference's DDP shape (batch / N rays per rank): after every step all ranks
    hold bit-identical parameters, and they equal a single-process Trainer on the whole batch up to summation order.
  * render_sharded: one frame split with shard_bounds, LR results gathered, equal to the single-GPU frame bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nerf_oracle as O

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from nerf_sr_b200 import Renderer, Trainer
        from nerf_sr_b200.parallel import P2PGradReducer, TorchGradReducer, render_sharded, shard_bounds
        res = {}
        cfg = O.RenderConfig(noise_std=1.0)
        pc, pf = O.make_mlp_params(cfg, 21), O.make_mlp_params(cfg, 8)
        r = Renderer(cfg, dev, precision="bf16x3")
        # ---- the all-reduce, both implementations, same data
        n = 2 * 595844
        p2p, ref = P2PGradReducer(r, n), TorchGradReducer(n, dev)
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        for it in range(3):
            x = torch.randn(n, device=dev, generator=g)
            p2p.buffer[:n].copy_(x)
            ref.buffer.copy_(x)
            p2p.allreduce_mean_()
            ref.allreduce_mean_()
            torch.cuda.synchronize()
            parts = [torch.empty(n, device=dev) for _ in range(world)]
            dist.all_gather(parts, x)
            want = parts[0].clone()
            for t in parts[1:]:
                want = want + t
            want = want * torch.tensor(1.0 / world, device=dev)
            res[f"p2p_exact_{it}"] = bool(torch.equal(p2p.buffer[:n], want))
            res[f"nccl_close_{it}"] = float((ref.buffer - want).abs().max())
        p2p.close()
        # ---- data-parallel training at the DDP shape: 512 rays in total
        s = 2
        rays_all = O.synthetic_rays(512, 300, "llff")
        tgt_all = torch.rand(128, 3, generator=torch.Generator().manual_seed(5))
        lo, hi = shard_bounds(512, world, s * s)[rank]
        tr = Trainer(r, pc, pf, downscale=s)
        gen = torch.Generator().manual_seed(77)                      # every rank draws the GLOBAL batch's randomness, takes its rows
        for step in range(3):
            full = {"u_coarse": torch.rand(512, 64, generator=gen), "noise_coarse": torch.randn(512, 64, generator=gen),
                    "u_fine": torch.rand(512, 64, generator=gen), "noise_fine": torch.randn(512, 128, generator=gen)}
            tr.optimize_parameters(rays_all[lo:hi].to(dev), tgt_all[lo // 4: hi // 4].to(dev), {k: v[lo:hi].to(dev) for k, v in full.items()})
        torch.cuda.synchronize()
        flat = torch.cat([p.reshape(-1) for w in (0, 1) for p in tr.params[w]])
        parts = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(parts, flat)
        res["params_identical_across_ranks"] = all(bool(torch.equal(parts[0], t)) for t in parts[1:])
        res["allreduce_impl"] = tr.allreduce_impl
        if rank == 0:                                                # the same 3 steps in ONE process on the whole batch
            r1 = Renderer(cfg, dev, precision="bf16x3")
            tr1 = Trainer.__new__(Trainer)
            Trainer.__init__(tr1, r1, pc, pf, downscale=s, allreduce="nccl")
            tr1._reducer = None                                       # single process: no collective
            tr1._gflat = torch.empty(2 * tr1._numel, device=dev)
            gen = torch.Generator().manual_seed(77)
            for step in range(3):
                full = {"u_coarse": torch.rand(512, 64, generator=gen), "noise_coarse": torch.randn(512, 64, generator=gen),
                        "u_fine": torch.rand(512, 64, generator=gen), "noise_fine": torch.randn(512, 128, generator=gen)}
                tr1.optimize_parameters(rays_all.to(dev), tgt_all.to(dev), {k: v.to(dev) for k, v in full.items()})
            flat1 = torch.cat([p.reshape(-1) for w in (0, 1) for p in tr1.params[w]])
            res["max_param_diff_vs_single_process_over_lr"] = float((flat - flat1).abs().max()) / 5e-4
            res["mean_param_diff_vs_single_process_over_lr"] = float((flat - flat1).abs().mean()) / 5e-4
            r1.close()
        # ---- one frame, ray-sharded, gathered
        cfg_e = O.RenderConfig(white_bkgd=True)
        re = Renderer(cfg_e, dev, precision="bf16x3")
        re.load_state_dict(0, O.make_mlp_params(cfg_e, 4))
        re.load_state_dict(1, O.make_mlp_params(cfg_e, 17))
        frame = O.synthetic_rays(4 * 1237, 9, "blender").to(dev)

        def render(shard):
            o = re.forward_rays(shard, want_weights=False)
            return re.box_average(o["fine_comp_rgbs"], s), re.box_average(o["fine_depth"], s)
        (rgb, depth), bounds = render_sharded(render, frame, s)
        rgb1, depth1 = render(frame)
        res["sharded_frame_equal"] = bool(torch.equal(rgb, rgb1) and torch.equal(depth, depth1))
        re.close()
        r.close()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_multi_gpu_paths(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, r in res.items():
        for it in range(3):
            assert r[f"p2p_exact_{it}"], (rank, r)
            assert r[f"nccl_close_{it}"] < 1e-5, (rank, r)
        assert r["params_identical_across_ranks"] and r["sharded_frame_equal"], (rank, r)
        assert "nsr_comm_allreduce_mean" in r["allreduce_impl"]
    # a sum of two half-batch means vs one full-batch mean differs only by fp32 summation order -> Adam moves parameters
    # by ~lr per step either way; the trajectories must not separate by more than a few lr, and not at all on average
    assert res[0]["max_param_diff_vs_single_process_over_lr"] <= 6.5 and res[0]["mean_param_diff_vs_single_process_over_lr"] < 0.3, res[0]
    import json
    from conftest import ROOT
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r02_multi_gpu.jsonl"), "a") as fh:
        fh.write(json.dumps({str(k): v for k, v in res.items()}) + "\n")
