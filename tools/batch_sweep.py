"""forward_rays latency / throughput versus batch size (the reference calls it per 4096-ray chunk;
a training step renders 2048 rays).  Usage: python tools/batch_sweep.py [precision]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from nerf_sr_b200 import Renderer

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
cfg, pc, pf, rays = bench.make_inputs()
r = Renderer(cfg, torch.device("cuda:0"), precision=prec)
r.load_state_dict(0, pc); r.load_state_dict(1, pf)
rays = rays.cuda()
for n in (1, 2, 128, 296, 1024, 2048, 4096, 16384, 65536, 160000):
    x = rays[:n].contiguous()
    for _ in range(3):
        r.forward_rays(x)
    torch.cuda.synchronize()
    reps = 20 if n <= 16384 else 5
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    for _ in range(reps):
        r.forward_rays(x)
    e.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    dev = s.elapsed_time(e) / reps
    print(f"n={n:7d}  device {dev:8.3f} ms  wall {wall*1e3:8.3f} ms  -> {n / (dev * 1e-3) / 1e6:7.3f} M rays/s", flush=True)
r.close()
