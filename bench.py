#!/usr/bin/env python
"""bench.py -- rays/sec of the NeRF-SR render hot path (64 coarse + 128 fine samples, 2x2 SS).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # the reference algorithm on host cores

A "step" renders one synthetic frame of BASELINE.json configs[1] (Blender-like 400x400 output
pixels, 2x2 super-sampling -> 160 000 HR rays, random-init 256-wide coarse+fine MLPs).  One JSON
line is printed by rank 0.  `value` times the device-resident path (rays already in HBM), `e2e`
the host-buffer call (pinned rays in, LR rgb+depth out, copies inside the timed region),
`roofline` the dominant kernel (fine pass) against the measured tensor peak, `cpu_baseline` the
oracle port (the reference's algorithm, torch CPU ops) on a bounded sample of the same workload.
The same line carries what BASELINE.json's other configs ask for, measured at every N the driver runs:
`strong` = ONE frame of configs[4] (762 048 rays) and of configs[3] (640 000 rays, 4x4 SS) held in rank 0's
pinned host memory, ray-sharded over the N ranks (H2D on rank 0, NCCL scatter, render, box average, NCCL
gather, D2H on rank 0 -- all inside the timed region); `train` = configs[2] at the reference's DDP shape
(2048 rays per step IN TOTAL, 2048 / N per rank, gradient all-reduce inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FLOP_PER_POINT = 1186816            # 2 x 593 408 MAC  
N_COARSE, N_IMPORTANCE = 64, 64
WORKLOAD = "blender-like 400x400 LR pixels... see config"
LR_W = LR_H = 200                   # configs[1]: 400x400 HR rays, downscale 2 -> 200x200 LR pixels
SS = 2
RAYS_PER_FRAME = LR_W * LR_H * SS * SS    # 160 000
SEEDS = (4, 17)                     # non-degenerate kaiming seeds (tests/golden uses the same)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback"


def ncu_traffic(kernel: str, field: str):
    """profiles/r02_ncu_traffic.json (written from this round's `ncu --set full` capture by tools/ncu_summary.py):
    {kernel: {mean_bytes_per_launch, fine_bytes_per_launch, source}}.  None when no capture of this round is committed."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    try:
        return json.load(open(p)).get(kernel, {}).get(field)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML in-process (a sample
    every ~5 ms, so even a 100 ms timed region is covered); falls back to the nvidia-smi query loop of the recipe."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows, self.how = index, False, [], None

    def _run_nvml(self) -> bool:
        try:
            import pynvml as N
            N.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].strip().isdigit() else self.index
            h = N.nvmlDeviceGetHandleByIndex(phys)
            mx = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            reasons_fn = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        except Exception:
            return False
        self.how = "nvml, ~5 ms period"
        while not self.stop_flag:
            try:
                r = int(reasons_fn(h))
                self.rows.append([float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), mx, N.nvmlDeviceGetPowerUsage(h) / 1e3] +
                                 [bool(r & bits[n]) for n in self.NAMES])
            except Exception:
                pass
            time.sleep(0.005)
        return True

    def run(self):
        if self._run_nvml():
            return
        self.how = "nvidia-smi, 200 ms period"
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append([float(f[0]), float(f[1]), float(f[2])] + [x.lower().startswith("active") for x in f[3:7]])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        pw = sorted(r[2] for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[3 + i] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.rows[0][1], "power_w_median": pw[len(pw) // 2],
                "power_w_max": pw[-1], "reasons": reasons, "samples": len(self.rows), "how": self.how}


def make_inputs(seed_offset: int = 0):
    # input generation only (seeded rays + kaiming weights); no oracle code on the GPU arm
    from nerf_sr_b200 import synthetic as S
    cfg = S.RenderConfig(white_bkgd=True, N_coarse=N_COARSE, N_importance=N_IMPORTANCE, downscale=SS)
    pc, pf = S.make_mlp_params(cfg, SEEDS[0]), S.make_mlp_params(cfg, SEEDS[1])
    rays = S.synthetic_rays(RAYS_PER_FRAME, 100 + seed_offset, "blender")
    return cfg, pc, pf, rays


def cpu_reference_rate(cfg, pc, pf, rays, sample_rays: int, repeats: int):
    """The reference algorithm (oracle port, torch CPU ops), eval mode, chunk 4096, on the host's
    cores.  The thread count is calibrated (all cores vs fewer: small GEMMs oversubscribe badly on
    100+ core hosts) so the baseline is the best the reference's code path does on this box."""
    from oracle import nerf_oracle as O
    ncpu = os.cpu_count() or 1
    sample = rays[:sample_rays]
    with torch.no_grad():
        best = None
        for nt in sorted({ncpu, max(1, ncpu // 2), min(ncpu, 32), min(ncpu, 16)}, reverse=True):
            torch.set_num_threads(nt)
            O.chunked_forward(pc, pf, sample[:512], cfg)   # warm-up
            t0 = time.perf_counter()
            O.chunked_forward(pc, pf, sample[:1024], cfg)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, nt)
        torch.set_num_threads(best[1])
        cpu_reference_rate.threads = best[1]
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.chunked_forward(pc, pf, sample, cfg)
            times.append(time.perf_counter() - t0)
    times.sort()
    return sample_rays / times[len(times) // 2], times


def torch_gpu_reference_rate(cfg, pc, pf, rays_dev, n_chunks: int = 4, chunk: int = 4096):
    """the reference's PyTorch path on the SAME GPU (oracle port = the reference's
    ATen op sequence, fp32, allow_tf32 off, ray_chunk 4096) -- the like-for-like 'before' number."""
    from oracle import nerf_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = rays_dev.device
    pcd = {k: v.to(dev) for k, v in pc.items()}
    pfd = {k: v.to(dev) for k, v in pf.items()}
    sample = rays_dev[: n_chunks * chunk]
    with torch.no_grad():
        O.chunked_forward(pcd, pfd, sample[:chunk], cfg, chunk)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        O.chunked_forward(pcd, pfd, sample, cfg, chunk)
        b.record()
        torch.cuda.synchronize()
    return sample.shape[0] / (a.elapsed_time(b) * 1e-3)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg, pc, pf, rays = make_inputs()
    sample = 4096
    for _ in range(args.warmup):
        pass
    rate, times = cpu_reference_rate(cfg, pc, pf, rays, sample, max(1, args.steps))
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": "rays/sec (64+128 samples, 2x SS)", "value": rate, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sorted(times)[len(times) // 2], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": getattr(cpu_reference_rate, "threads", cores), "kind": "port",
                         "host_cores": cores,
                         "sample": f"{sample} rays of the frame per step (chunk 4096), median of {len(times)}; "
                                   "thread count calibrated over {all, 1/2, 32, 16}"},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "BASELINE configs[1]: Blender-like 400x400 HR rays (200x200 LR pixels x 2x2 SS), "
                        "64 coarse + 128 fine samples, random-init 256-wide coarse+fine MLP, eval mode",
            "rays_per_step_per_gpu": RAYS_PER_FRAME, "n_coarse": N_COARSE, "n_fine": N_COARSE + N_IMPORTANCE,
            "supersampling": SS, "parallelism": f"ray-sharded x{n_gpus} (independent frames, no collective)",
            "l2": "flushed between timed steps (256 MiB write)"}


# ------------------------------------------------------------------------------------------------
# --workload train: BASELINE configs[2] shape (LLFF-like rays, 512 LR pixels x 2x2 SS = 2048 rays per rank per
# step, 64 + 128 samples, sigma noise 1.0): forward (train mode) + LR loss + backward + gradient all-reduce (N > 1)
# + Adam + weight re-pack, the reference's optimize_parameters (models/nerf_downX_model.py:398-408).
# ------------------------------------------------------------------------------------------------
TRAIN_N_LR = 512


def train_inputs(rank: int):
    from nerf_sr_b200 import synthetic as S
    cfg = S.RenderConfig(noise_std=1.0, N_coarse=N_COARSE, N_importance=N_IMPORTANCE, downscale=SS)
    pc, pf = S.make_mlp_params(cfg, 21), S.make_mlp_params(cfg, 8)
    rays = S.synthetic_rays(TRAIN_N_LR * SS * SS, 300 + rank, "llff")
    target = torch.rand(TRAIN_N_LR, 3, generator=torch.Generator().manual_seed(500 + rank))
    return cfg, pc, pf, rays, target


def oracle_train_rate(cfg, pc, pf, rays, target, device, n_lr: int, repeats: int):
    """The reference's training iteration (oracle port: torch autograd + Adam restatement) on `device`."""
    from oracle import nerf_oracle as O
    from oracle import train_oracle as T
    dev = torch.device(device)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    st = T.TrainState({k: v.to(dev) for k, v in pc.items()}, {k: v.to(dev) for k, v in pf.items()})
    r, t = rays[: n_lr * SS * SS].to(dev), target[:n_lr].to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    tc = T.TrainConfig()

    def draw():
        n = r.shape[0]
        return O.RenderRng(torch.rand(n, cfg.N_coarse, device=dev, generator=g), torch.randn(n, cfg.N_coarse, device=dev, generator=g),
                           torch.rand(n, cfg.N_importance, device=dev, generator=g),
                           torch.randn(n, cfg.N_coarse + cfg.N_importance, device=dev, generator=g))
    T.optimize_parameters(st, r, t, cfg, tc, draw(), SS)
    times = []
    for _ in range(repeats):
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        T.optimize_parameters(st, r, t, cfg, tc, draw(), SS)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    times.sort()
    return r.shape[0] / times[len(times) // 2]


def train_config(world):
    return {"workload": "BASELINE configs[2] shape: LLFF-like rays, 512 LR pixels x 2x2 SS = 2048 rays per GPU per step, 64 coarse + "
                        "128 fine samples, sigma noise 1.0, random-init 256-wide coarse+fine MLP; one step = forward (train mode) + "
                        "LR MSE loss + backward + Adam + weight re-pack",
            "rays_per_step_per_gpu": TRAIN_N_LR * SS * SS, "n_coarse": N_COARSE, "n_fine": N_COARSE + N_IMPORTANCE, "supersampling": SS,
            "parallelism": f"ray-sharded x{world}" + (", NCCL all-reduce (mean) of the 4.77 MB gradient bucket per step" if world > 1 else ""),
            "l2": "working set per step (5 GB activation stash) exceeds L2"}


def run_reference_train(args, rank):
    """--impl reference --workload train: the reference's training iteration (oracle port) on the host cores."""
    if rank != 0:
        return
    cfg, pc, pf, rays, target = train_inputs(0)
    ncpu = os.cpu_count() or 1
    torch.set_num_threads(min(ncpu, 32))
    n_lr = 64
    rate = oracle_train_rate(cfg, pc, pf, rays, target, "cpu", n_lr, max(1, args.steps))
    line = {"impl": "reference", "metric": "training rays/sec (64+128 samples, 2x SS; forward + backward + Adam)", "value": rate,
            "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_lr * SS * SS / rate,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": train_config(1),
            "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": min(ncpu, 32), "host_cores": ncpu, "kind": "port",
                             "sample": f"{n_lr * SS * SS} rays ({n_lr} LR pixels) per step, median of {max(1, args.steps)} iterations of the oracle "
                                       "port (torch CPU autograd + Adam)"},
            "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_train(args, rank, world, local):
    import torch.distributed as dist
    from nerf_sr_b200 import Renderer, Trainer
    dev = torch.device(f"cuda:{local}")
    cfg, pc, pf, rays_cpu, target_cpu = train_inputs(rank)
    r = Renderer(cfg, dev, precision="bf16x3")
    tr = Trainer(r, pc, pf, downscale=SS)
    rays, target = rays_cpu.to(dev), target_cpu.to(dev)
    rays_pin, target_pin = rays_cpu.pin_memory(), target_cpu.pin_memory()
    gen = torch.Generator(device=dev).manual_seed(rank)
    n = rays.shape[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        tr.optimize_parameters(rays, target, tr.draw_rng(n, gen))
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    launches0 = r.launch_count
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        tr.optimize_parameters(rays, target, tr.draw_rng(n, gen))
    b.record()
    barrier()
    dev_ms = a.elapsed_time(b)
    launches = r.launch_count - launches0
    # end to end: the batch comes from pinned host memory every step, the step's losses go back to the host
    dev_rays, dev_tgt = torch.empty_like(rays), torch.empty_like(target)
    host_metrics = torch.empty(4, dtype=torch.float32).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dev_rays.copy_(rays_pin, non_blocking=True)
        dev_tgt.copy_(target_pin, non_blocking=True)
        m = tr.optimize_parameters(dev_rays, dev_tgt, tr.draw_rng(n, gen))
        host_metrics.copy_(m, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True
    sampler.join(timeout=2)
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = (float(x) for x in t.tolist())
    if rank == 0:
        pk, pk_kind = peaks()
        total = n * world * args.steps
        value = total / (dev_ms / 1e3)
        line = {
            "metric": "training rays/sec (64+128 samples, 2x SS; forward + backward + Adam)", "value": value, "unit": "rays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3 split (fp32 accumulate, fp32 master weights)",
            "data": "synthetic", "config": train_config(world),
            "e2e": {"value": total / (e2e_ms / 1e3), "unit": "rays/s", "h2d_bytes_per_step": n * 8 * 4 + (n // (SS * SS)) * 3 * 4,
                    "d2h_bytes_per_step": 16, "api": "Trainer.optimize_parameters (pinned host batch in, losses out)"},
            "gpu_launches": int(launches),
            "final_loss": [float(x) for x in host_metrics.tolist()],
            "roofline": train_roofline(n, args.steps, dev_ms, pk, pk_kind),
            "clocks": sampler.summary(),
        }
        if world == 1 and args.torch_gpu_port:
            try:
                line["torch_gpu_reference_port"] = {
                    "value": oracle_train_rate(cfg, pc, pf, rays_cpu, target_cpu, f"cuda:{local}", TRAIN_N_LR, 3), "unit": "rays/s",
                    "what": "oracle port of optimize_parameters (the reference's PyTorch autograd + Adam op sequence) on this GPU, "
                            "fp32, allow_tf32 off, 2048 rays per step; not the product path"}
            except Exception as e:
                line["torch_gpu_reference_port"] = {"unavailable": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            ncpu = os.cpu_count() or 1
            torch.set_num_threads(min(ncpu, 32))
            rate = oracle_train_rate(cfg, pc, pf, rays_cpu, target_cpu, "cpu", 64, 3)
            line["cpu_baseline"] = {"value": rate, "unit": "rays/s", "cores": min(ncpu, 32), "host_cores": ncpu, "kind": "port",
                                    "sample": "256 rays (64 LR pixels) per step, median of 3 full iterations of the oracle port "
                                              "(torch CPU autograd + Adam)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    r.close()


# HBM bytes one point (one sample of one ray) costs in a training iteration, by design (every operand of
# the backward GEMMs is a bf16 hi/lo "tile image" written once and read once):
#   forward stash written   enc 256 + h_1..h_8, feat 9 x 1024 + dir 512 + ReLU bits 256 + raw 16 + z 4            = 10 260
#   k_render_bwd            reads raw 16, z 4, dir 512, noise 4; writes dHead 256, dZ_dir 512, encdir 256, d sigma 4 =  1 564
#   k_tg_dxchain            reads dZ_dir 512, ReLU bits 256, d sigma 4; writes the nine dZ images 9 x 1024           =  9 988
#   k_tg_dw (13 GEMMs)      both operands once: rgb 768, dir 1536, dir-enc 768, final+sigma 2304, L8,7,6,4,3,2 6 x 2048,
#                           L5 2048, L5-enc 1280, L1 1280                                                            = 22 272
TRAIN_BYTES_PER_POINT = 10260 + 1564 + 9988 + 22272

# The dW GEMMs are split over the SMs along the points; every split writes an fp32 partial [N][128] (+ 128 bias sums) that
# k_grad_reduce reads back in a fixed order (deterministic, no atomics).  This traffic does not scale with the batch:
# (jobs, N) of the 13 GEMMs of a net (csrc/nsr_train.cu backward_net): rgb, dir, dir-enc, final+sigma, L8 L7 L6 L5 L5-enc L4 L3 L2, L1
DW_GEMMS = [(1, 128), (1, 256), (1, 64), (3, 256)] + [(2, 256)] * 3 + [(2, 256), (2, 64)] + [(2, 256)] * 3 + [(2, 64)]


def dw_partial_bytes(n_rays: int, sm_count: int = 148) -> int:
    """Bytes of dW partials written + read back per training step (both nets), mirroring DwPlanner::launch."""
    total = 0
    for s in (N_COARSE, N_COARSE + N_IMPORTANCE):
        tiles = (n_rays * s + 127) // 128
        for jobs, n in DW_GEMMS:
            n_split = max(1, min(sm_count // jobs, (tiles + 3) // 4))
            total += 2 * jobs * n_split * 128 * (n + 1) * 4
    return total


def train_roofline(n: int, steps: int, dev_ms: float, pk: dict, pk_kind: str) -> dict:
    """The `roofline` object of the training line (pure: unit-tested on CPU).  The iteration is HBM-bound by construction,
    so the bound is the measured copy bandwidth: achieved = bytes the step's design moves (TRAIN_BYTES_PER_POINT x points,
    PLUS the dW partial sums written and read back, which do not scale with the batch) / event-timed step time.  The
    tensor-side view (algorithmic FLOP = 3 x forward) is kept beside it.  n: rays per step per GPU; dev_ms: all steps."""
    points = n * (N_COARSE + N_COARSE + N_IMPORTANCE)
    partial_bytes = dw_partial_bytes(n)
    step_bytes = points * TRAIN_BYTES_PER_POINT + partial_bytes
    gbs = step_bytes * steps / (dev_ms / 1e3) / 1e9
    flop_step = points * FLOP_PER_POINT * 3                     # forward + 2x backward
    tflops = flop_step * steps / (dev_ms / 1e3) / 1e12
    tpeak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    return {
        "bound": "hbm", "kernel": "whole step (k_tc_pass stash variant, k_render_bwd, k_tg_dxchain, k_tg_dw; DESIGN.md section 11)",
        "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
        "peak_kind": f"{pk_kind} copy bandwidth", "bytes_per_step": step_bytes, "bytes_per_point": TRAIN_BYTES_PER_POINT,
        "dw_partial_bytes_per_step": partial_bytes,
        "traffic": None,
        "tensor": {"achieved": tflops, "peak": tpeak, "unit": "TFLOP/s", "frac": tflops / tpeak, "issued_frac": 3 * tflops / tpeak,
                   "flop_per_step": flop_step, "peak_kind": f"{pk_kind} cuBLAS bf16 sustained"},
    }


def render_roofline(precision: str, n: int, steps: int, dev_ms: float, fine_ms: float, pk: dict, pk_kind: str,
                    launches_per_step: float = 1.0) -> dict:
    """The `roofline` object of the render line (pure: unit-tested on CPU).

    Dominant kernel = k_tc_pass.  On the tensor-core path a step is ONE launch of its frame variant (coarse S=64 tiles, the
    resampling, fine S=128 tiles and the box average; profiles/r02_launch_list.md) -- `launches_per_step` is what the
    library counted in the timed region (6 with NSR_TC_FUSED=0: coarse, fine, 4 box averages; the fp32 path has more).
    `achieved` = the algorithmic FLOP of the step (1 186 816 FLOP per point x (64 + 128) points per ray) over
    the step time measured with CUDA events INSIDE the timed region, so the denominator is the SUSTAINED measured peak (a
    kernel timed inside a long step).  The fine pass timed ALONE as its own launch (nsr_render_pass; its own event pair) is
    reported next to it, against both the sustained and the burst peak.
    n: rays per step per GPU; dev_ms: device time of all `steps` steps (max over ranks); fine_ms: one fine-pass launch."""
    step_flops = n * (N_COARSE + N_COARSE + N_IMPORTANCE) * FLOP_PER_POINT
    achieved = step_flops * steps / (dev_ms / 1e3) / 1e12
    fine_flops = n * (N_COARSE + N_IMPORTANCE) * FLOP_PER_POINT
    fine_achieved = fine_flops / (fine_ms / 1e3) / 1e12
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    split = 3 if "x3" in precision else 1
    tc = precision != "fp32_simt"
    return {
        "bound": "tensor", "kernel": ("k_tc_pass, frame variant: coarse S=64 + fine S=128 tiles of one frame in one launch" if launches_per_step == 1
                                      else "k_tc_pass (coarse S=64 + fine S=128 launches of one step)") if tc else "k_simt_mlp",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "peak_kind": f"{pk_kind} cuBLAS bf16 sustained (kernel timed inside the timed region of back-to-back steps)",
        "issued_frac": split * achieved / peak,
        "flop_per_launch": step_flops / (1 if launches_per_step == 1 else 2),
        "ms_per_launch": dev_ms / steps / (1 if launches_per_step == 1 else 2),
        "launches_per_step": launches_per_step,
        # dram__bytes_read.sum + dram__bytes_write.sum per launch: NOT measurable inside this process (it needs ncu
        # replays); taken from this round's committed `ncu --set full` capture when it exists, with its provenance
        "traffic": ncu_traffic("k_tc_pass", "frame_bytes_per_launch" if launches_per_step == 1 else "mean_bytes_per_launch") if tc else None,
        "traffic_source": ncu_traffic("k_tc_pass", "source") if tc else None,
        # frame variant, per ray: 32 B rays in + 2 x 20 B HR results + 2 x 16 B / s^2 LR results out (the 512 B of fine
        # z-values per ray are an L2-resident intermediate); separate launches: 32 B + 512 B + 20 B per ray and launch
        "algorithmic_bytes_per_launch": n * (32 + 40 + 32 // (SS * SS)) if launches_per_step == 1 else n * (32 + 4 * (N_COARSE + N_IMPORTANCE) + 20),
        "fine_pass_alone": {"achieved": fine_achieved, "flop_per_launch": fine_flops, "ms_per_launch": fine_ms,
                            "frac_vs_sustained_peak": fine_achieved / peak, "burst_peak": pk["bf16_tflops"],
                            "frac_vs_burst_peak": fine_achieved / pk["bf16_tflops"],
                            "traffic": ncu_traffic("k_tc_pass", "fine_bytes_per_launch") if tc else None},
    }



# ------------------------------------------------------------------------------------------------
# `strong`: one frame, ray-sharded over the ranks (BASELINE configs[3] / configs[4])
# ------------------------------------------------------------------------------------------------
STRONG_FRAMES = {
    # name: (rays, s, kind, white_bkgd)
    "C5_llff_1008x756_s2": (1008 * 756, 2, "llff", False),
    "C4_blender_800x800_s4": (800 * 800, 4, "blender", True),
}


def run_strong(r_by_cfg, rank, world, dev, steps, warmup):
    """One frame per step, held in rank 0's PINNED HOST memory; per step and inside the timed region: rank 0 uploads the
    frame (H2D), NCCL scatters contiguous ray shards (boundaries at multiples of s*s: parallel.shard_bounds), every
    rank renders its shard (no collective in the render) and box-averages it, NCCL gathers LR rgb+depth on rank 0,
    rank 0 copies the LR frame to pinned host memory (D2H).  Device-timed with CUDA events on every rank's stream,
    max over ranks.  Returns {frame: {...}} on every rank (identical after the MAX all-reduce)."""
    import torch.distributed as dist
    from nerf_sr_b200 import synthetic as S
    from nerf_sr_b200.parallel import shard_bounds
    out = {}
    for name, (n, s, kind, white) in STRONG_FRAMES.items():
        r = r_by_cfg(white, s)
        ss = s * s
        bounds = shard_bounds(n, world, ss)
        lo, hi = bounds[rank]
        mx = max(b[1] - b[0] for b in bounds)
        if rank == 0:
            frame_pin = S.synthetic_rays(n, 4242, kind).pin_memory()
            frame_dev = torch.empty(world * mx, 8, device=dev)             # shard-padded layout for the scatter
            lr_pin = torch.empty(n // ss, 4).pin_memory()
            gathered = [torch.empty(mx // ss, 4, device=dev) for _ in range(world)]
        shard = torch.empty(mx, 8, device=dev)
        lr = torch.empty(mx // ss, 4, device=dev)

        def one_frame():
            if rank == 0:
                if world == 1:
                    shard.copy_(frame_pin, non_blocking=True)
                else:
                    for g, (a, b) in enumerate(bounds):                          # H2D straight into the scatter layout
                        frame_dev[g * mx: g * mx + (b - a)].copy_(frame_pin[a:b], non_blocking=True)
            if world > 1:
                dist.scatter(shard, list(frame_dev.view(world, mx, 8).unbind(0)) if rank == 0 else None, src=0)
            o = r.forward_rays(shard[: hi - lo], want_weights=False)
            lr[: (hi - lo) // ss, :3] = r.box_average(o["fine_comp_rgbs"], s)
            lr[: (hi - lo) // ss, 3:] = r.box_average(o["fine_depth"], s)
            if world > 1:
                dist.gather(lr, gathered if rank == 0 else None, dst=0)
                if rank == 0:
                    for g, (a, b) in enumerate(bounds):
                        lr_pin[a // ss: b // ss].copy_(gathered[g][: (b - a) // ss], non_blocking=True)
            elif rank == 0:
                lr_pin.copy_(lr[: n // ss], non_blocking=True)

        for _ in range(max(1, warmup // 2)):
            one_frame()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(steps):
            one_frame()
        b.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        t = torch.tensor([a.elapsed_time(b), wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, wall_ms = float(t[0]) / steps, float(t[1]) / steps
        if rank == 0:
            assert bool(torch.isfinite(lr_pin).all())
        out[name] = {"rays": n, "s": s, "rays_per_rank": [b_ - a_ for a_, b_ in bounds], "ms_per_frame": ms,
                     "wall_ms_per_frame": wall_ms, "value": n / (ms / 1e3), "unit": "rays/s",
                     "h2d_bytes_per_frame": n * 32, "d2h_bytes_per_frame": (n // ss) * 16,
                     "nccl_bytes_per_frame": 0 if world == 1 else (world - 1) * (mx * 32 + (mx // ss) * 16),
                     "timed": "CUDA events on each rank's stream around K whole frames (H2D on rank 0, scatter, render, box "
                              "average, gather, D2H on rank 0), max over ranks"}
    return out


def run_train_ddp(rank, world, dev, steps, warmup):
    """BASELINE configs[2] at the reference's DDP shape (data/__init__.py:94-99: per-rank batch = batch_size / n_gpus):
    512 LR pixels x 2x2 SS = 2048 rays per step IN TOTAL, 2048 / N per rank; forward (train mode) + LR loss + backward +
    gradient all-reduce (mean over ranks) + Adam + re-pack, device-timed with the all-reduce inside the region."""
    import torch.distributed as dist
    from nerf_sr_b200 import Renderer, Trainer
    from nerf_sr_b200.parallel import shard_bounds
    cfg, pc, pf, rays_cpu, target_cpu = train_inputs(0)                   # every rank derives the same global batch
    lo, hi = shard_bounds(rays_cpu.shape[0], world, SS * SS)[rank]
    rays, target = rays_cpu[lo:hi].to(dev), target_cpu[lo // (SS * SS): hi // (SS * SS)].to(dev)
    r = Renderer(cfg, dev, precision="bf16x3")
    tr = Trainer(r, pc, pf, downscale=SS)
    gen = torch.Generator(device=dev).manual_seed(rank)
    n = rays.shape[0]
    for _ in range(max(3, warmup)):
        tr.optimize_parameters(rays, target, tr.draw_rng(n, gen))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = r.launch_count
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        tr.optimize_parameters(rays, target, tr.draw_rng(n, gen))
    b.record()
    torch.cuda.synchronize()
    launches = r.launch_count - launches0
    # the collective alone, same buffers, same stream (event-timed): what it adds to the step
    ar_ms = 0.0
    if world > 1:
        gc, gf = tr.last_grads
        c, d = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        c.record()
        for _ in range(20):
            tr.allreduce_grads(gc, gf)
        d.record()
        torch.cuda.synchronize()
        ar_ms = c.elapsed_time(d) / 20
    t = torch.tensor([a.elapsed_time(b) / steps, ar_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ar_ms = float(t[0]), float(t[1])
    res = {"rays_per_step_total": rays_cpu.shape[0], "rays_per_step_per_gpu": n, "ms_per_step": ms,
           "value": rays_cpu.shape[0] / (ms / 1e3), "unit": "rays/s", "scaling": "strong",
           "allreduce_ms": ar_ms, "allreduce_bytes": 2 * int(r.lib.nsr_grad_numel(r._h)) * 4,
           "allreduce_impl": tr.allreduce_impl if world > 1 else None,
           "gpu_launches_per_step": launches / steps,
           "final_loss": [float(x) for x in tr.last_metrics.tolist()],
           "timed": "CUDA events around K optimize_parameters calls (gradient all-reduce inside), max over ranks"}
    r.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16x3", choices=["fp32_simt", "bf16x3", "fp16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-gpu-port", dest="torch_gpu_port", action="store_false",
                    help="skip timing the reference's PyTorch op sequence (oracle port, fp32, TF32 off) on this GPU -- the "
                         "same-GPU 'before' number (BASELINE.md section 3), measured by default on rank 0 at N=1 (~0.3 s)")
    ap.add_argument("--no-extras", action="store_true", help="skip the `strong` / `train` / `e2e_full` / weights-variant sub-objects")
    ap.set_defaults(torch_gpu_port=True)
    ap.add_argument("--workload", default="render", choices=["render", "train"],
                    help="render = the headline metric (BASELINE configs[1]); train = the training iteration (configs[2] shape)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.workload == "train":
            run_reference_train(args, rank)
        else:
            run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from nerf_sr_b200 import Renderer
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no GPU visible; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 must print exactly ONE line on stdout, but NCCL / c10d write a version banner there when
        # the communicator is created: send fd 1 to stderr while the group and its first collective
        # are set up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    if args.workload == "train":
        run_train(args, rank, world, local)
        return

    cfg, pc, pf, rays_cpu = make_inputs(seed_offset=rank)
    r = Renderer(cfg, dev, precision=args.precision)
    r.load_state_dict(0, pc)
    r.load_state_dict(1, pf)
    rays = rays_cpu.to(dev)
    rays_pinned = rays_cpu.pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n = rays.shape[0]

    def step_device():
        # forward() + comp_low_res_output of one frame (models/nerf_downX_model.py:316-348): HR composites / depths / opacities
        # and the LR (box-averaged) images of both nets -- ONE kernel launch on the tensor-core path (nsr_render_frame)
        return r.render_frame(rays, SS)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, flush_l2=True):
        """K calls of fn, each between its own CUDA event pair on the launching stream (L2 flushed in between)."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        for s_, e_ in ev:
            if flush_l2:
                flush.fill_(1)
            s_.record()
            fn()
            e_.record()
        torch.cuda.synchronize()
        return sum(s_.elapsed_time(e_) for s_, e_ in ev) / steps

    # what the reference's test() consumes per frame (models/nerf_downX_model.py:326-353,418-450): HR `_ori` rgb + depth and
    # their LR box averages, for the coarse and the fine net -- through the public Renderer API, pinned host buffers
    n_lr = n // (SS * SS)
    full_pin = {f"{net}_{k}": torch.empty(rows, c).pin_memory() for net in ("coarse", "fine")
                for k, rows, c in (("rgb_ori", n, 3), ("depth_ori", n, 1), ("rgb", n_lr, 3), ("depth", n_lr, 1))}
    rays_stage = torch.empty_like(rays)

    def step_e2e_full():
        rays_stage.copy_(rays_pinned, non_blocking=True)
        o = r.render_frame(rays_stage, SS)
        for net in ("coarse", "fine"):
            full_pin[f"{net}_rgb_ori"].copy_(o[f"{net}_comp_rgbs"], non_blocking=True)
            full_pin[f"{net}_depth_ori"].copy_(o[f"{net}_depth"].view(-1, 1), non_blocking=True)
            full_pin[f"{net}_rgb"].copy_(o[f"{net}_lr_rgb"], non_blocking=True)
            full_pin[f"{net}_depth"].copy_(o[f"{net}_lr_depth"].view(-1, 1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(args.warmup):
        step_device()
        r.render_frame_host(rays_pinned, SS)
    if not args.no_extras:
        step_e2e_full()
        r.forward_rays(rays, want_weights=True)
    torch.cuda.synchronize()

    # ---- device-resident throughput -----------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    launches0 = r.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for s, e in ev:
        flush.fill_(1)                  # evict L2 between timed steps (untimed)
        s.record()
        step_device()
        e.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = r.launch_count - launches0
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)

    # ---- dominant kernel (fine pass, one launch) for the roofline ------------------------
    out = r.forward_rays(rays, want_weights=False, want_z_fine=True)
    z_fine = out["z_fine"]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for s, e in kev:
        flush.fill_(1)
        s.record()
        r.render_pass(1, rays, z_fine)
        e.record()
    torch.cuda.synchronize()
    fine_ms = sum(s.elapsed_time(e) for s, e in kev) / args.steps

    sm_mhz_in_kernel = r.kernel_clock_mhz()      # clock64 / globaltimer stamps of the last fine-pass launch (CTA 0)

    # ---- end to end through the host-buffer call -----------------------------------------
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rgb, depth = r.render_frame_host(rays_pinned, SS)
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- variants: the full 8-key dict (weights maps included), and the e2e call returning what test() consumes ----
    ww_ms = full_s = 0.0
    if not args.no_extras:
        ww_ms = timed(lambda: r.forward_rays(rays, want_weights=True), args.steps)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e_full()
        barrier()
        full_s = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)

    t = torch.tensor([dev_ms, e2e_s * 1e3, fine_ms, ww_ms, full_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, fine_ms, ww_ms, full_ms = (float(x) for x in t.tolist())

    # ---- BASELINE configs[3] / [4] (one frame sharded over the ranks) and configs[2] at the DDP shape ----
    strong = train_ddp = None
    if not args.no_extras and args.precision == "bf16x3":
        cache = {}

        def r_by_cfg(white, s):
            from nerf_sr_b200 import synthetic as S
            if (white, s) not in cache:
                c2 = S.RenderConfig(white_bkgd=white, N_coarse=N_COARSE, N_importance=N_IMPORTANCE, downscale=s)
                rr = Renderer(c2, dev, precision=args.precision)
                rr.load_state_dict(0, pc)
                rr.load_state_dict(1, pf)
                cache[(white, s)] = rr
            return cache[(white, s)]
        strong = run_strong(r_by_cfg, rank, world, dev, max(2, min(args.steps, 5)), args.warmup)
        for rr in cache.values():
            rr.close()
        train_ddp = run_train_ddp(rank, world, dev, max(10, args.steps), args.warmup)

    if rank == 0:
        pk, pk_kind = peaks()
        total_rays = n * world * args.steps
        value = total_rays / (dev_ms / 1e3)
        line = {
            "metric": "rays/sec (64+128 samples, 2x SS)", "value": value, "unit": "rays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16x3": "bf16x3 split (fp32 accumulate)", "fp16x3": "fp16x3 split (fp32 accumulate)",
                      "fp32_simt": "f32", "bf16": "bf16"}[args.precision],
            "data": "synthetic", "config": workload_config(world),
            "lr_pixels_per_s": value / (SS * SS),
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
            "e2e": {"value": total_rays / (e2e_ms / 1e3), "unit": "rays/s", "h2d_bytes_per_step": n * rays.shape[1] * 4,
                    "d2h_bytes_per_step": (n // (SS * SS)) * 4 * 4, "api": "nsr_render_host (pinned host rays in, LR rgb+depth out)"},
            "gpu_launches": int(launches),
            "roofline": render_roofline(args.precision, n, args.steps, dev_ms, fine_ms, pk, pk_kind, launches / args.steps),
            "clocks": {**sampler.summary(), "sm_mhz_in_kernel": sm_mhz_in_kernel,
                       "sm_mhz_in_kernel_how": "clock64 / globaltimer deltas stamped by CTA 0 of the last fine-pass launch"},
        }
        if not args.no_extras:
            line["with_weights"] = {
                "value": n * world / (ww_ms / 1e3), "unit": "rays/s", "ms_per_step": ww_ms,
                "what": "forward_rays(want_weights=True): the reference's full 8-key dict incl. coarse_weights [N,64] and "
                        "fine_weights [N,128] (768 more bytes written per ray); no box average",
                "bytes_written_per_step": n * (2 * 20 + 4 * (N_COARSE + N_COARSE + N_IMPORTANCE))}
            line["e2e_full"] = {
                "value": total_rays / (full_ms / 1e3), "unit": "rays/s", "h2d_bytes_per_step": n * rays.shape[1] * 4,
                "d2h_bytes_per_step": 2 * (n * 16 + n_lr * 16),
                "api": "Renderer.render_frame (nsr_render_frame, one kernel launch per frame), pinned host rays in; HR `_ori` rgb+depth and LR rgb+depth of the "
                       "coarse AND the fine net out (what the reference's test() consumes, models/nerf_downX_model.py:326-353)"}
        if strong is not None:
            line["strong"] = strong
        if train_ddp is not None:
            line["train"] = train_ddp
        if world == 1 and args.torch_gpu_port:
            try:
                line["torch_gpu_reference_port"] = {
                    "value": torch_gpu_reference_rate(cfg, pc, pf, rays), "unit": "rays/s",
                    "what": "oracle port (the reference's PyTorch op sequence) on this GPU, fp32, allow_tf32 off, "
                            "ray_chunk 4096, 4 chunks; not the product path"}
            except Exception as e:      # never let the informational arm break the bench line
                line["torch_gpu_reference_port"] = {"unavailable": str(e)[:200]}
        if not args.no_cpu_baseline and world == 1:
            rate, times = cpu_reference_rate(cfg, pc, pf, rays_cpu, 4096, 3)
            line["cpu_baseline"] = {"value": rate, "unit": "rays/s", "cores": getattr(cpu_reference_rate, "threads", os.cpu_count() or 1),
                                    "host_cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": "4096 rays of the frame (one ray_chunk), median of 3, oracle port of the reference "
                                              "(torch CPU), thread count calibrated over {all, 1/2, 32, 16}",
                                    "note": "kind=port: oracle/nerf_oracle.py, a restatement pinned BIT-EQUAL to the unmodified "
                                            "reference in the build container (oracle/make_golden*.py); the reference tree itself "
                                            "is not a bench dependency"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    r.close()


if __name__ == "__main__":
    main()
