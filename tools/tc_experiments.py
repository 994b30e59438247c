"""Timing experiments on the fine pass (one launch, 160k rays): precision variants and debug flags."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from nerf_sr_b200 import Renderer

cfg, pc, pf, rays = bench.make_inputs()
rays = rays.cuda()
for prec, flags in (("bf16x3", 0), ("bf16x3", 1), ("bf16", 0), ("bf16", 1), ("fp16x3", 0)):
    r = Renderer(cfg, torch.device("cuda:0"), precision=prec)
    r.load_state_dict(0, pc); r.load_state_dict(1, pf)
    out = r.forward_rays(rays, want_weights=False, want_z_fine=True)
    z = out["z_fine"]
    r.lib.nsr_debug_set_flags(r._h, flags)
    for _ in range(2):
        r.render_pass(1, rays, z)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r.render_pass(1, rays, z); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    tiles_per_sm = 160000 / 148
    print(f"{prec:8s} flags={flags} fine pass {ms:7.2f} ms  -> {ms * 1e-3 * 1.9e9 / tiles_per_sm / 1e3:6.1f} kcycles/tile @1.9GHz "
          f"{160000 * 128 * 1186816 / ms / 1e9:7.1f} TFLOP/s algorithmic", flush=True)
    r.close()
