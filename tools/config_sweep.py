"""Single-GPU throughput at the sizes of BASELINE.json configs[0..4] (synthetic rays, eval mode):
C1 1024 rays coarse-only, C2 400x400 s=2, C4 800x800 s=4, C5 1008x756 s=2 (C3 is the training step:
forward only here, 2048 rays).  Device-resident and host-buffer (nsr_render_host) rates."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nerf_sr_b200 import Renderer
from nerf_sr_b200 import synthetic as O      # input generation only (seeded rays / kaiming weights)

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
dev = torch.device("cuda:0")
for name, n, s, nimp in (("C1 1024 rays, 64 coarse only", 1024, 1, 0), ("C3 fwd 2048 rays (512 LR px x 2x2)", 2048, 2, 64),
                         ("C2 400x400 s=2", 160000, 2, 64), ("C4 800x800 s=4", 640000, 4, 64), ("C5 1008x756 s=2", 762048, 2, 64)):
    cfg = O.RenderConfig(white_bkgd=True, N_importance=nimp)
    r = Renderer(cfg, dev, precision=prec)
    r.load_state_dict(0, O.make_mlp_params(cfg, 4)); r.load_state_dict(1, O.make_mlp_params(cfg, 17))
    rays_cpu = O.synthetic_rays(n, 7, "blender")
    rays = rays_cpu.cuda(); pinned = rays_cpu.pin_memory()
    for _ in range(3):
        r.forward_rays(rays, want_weights=False)
    r.render_frame_host(pinned, s)
    torch.cuda.synchronize()
    reps = 20 if n <= 4096 else 3
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        r.forward_rays(rays, want_weights=False)
    b.record(); torch.cuda.synchronize()
    dev_ms = a.elapsed_time(b) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        r.render_frame_host(pinned, s)
    host_ms = (time.perf_counter() - t0) / reps * 1e3
    print(f"{name:38s} device {dev_ms:9.3f} ms = {n / dev_ms / 1e3:6.3f} M rays/s | host in/out {host_ms:9.3f} ms = {n / host_ms / 1e3:6.3f} M rays/s", flush=True)
    r.close()
