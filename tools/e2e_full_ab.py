"""Where the time of bench.py's `e2e_full` step goes: the frame call alone, the copies alone, and both, for the one-launch
frame call and for forward_rays + box averages (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nerf_sr_b200 import Renderer

cfg, pc, pf, rays_cpu = bench.make_inputs()
dev = torch.device("cuda:0")
r = Renderer(cfg, dev, precision="bf16x3")
r.load_state_dict(0, pc); r.load_state_dict(1, pf)
SS = bench.SS
n = rays_cpu.shape[0]; n_lr = n // (SS * SS)
rays_pinned = rays_cpu.pin_memory(); rays_stage = torch.empty_like(rays_cpu, device=dev)
pin = {f"{net}_{k}": torch.empty(rows, c).pin_memory() for net in ("coarse", "fine")
       for k, rows, c in (("rgb_ori", n, 3), ("depth_ori", n, 1), ("rgb", n_lr, 3), ("depth", n_lr, 1))}

def frame_new():
    return r.render_frame(rays_stage, SS)
def frame_old():
    o = r.forward_rays(rays_stage, want_weights=False)
    for net in ("coarse", "fine"):
        o[f"{net}_lr_rgb"] = r.box_average(o[f"{net}_comp_rgbs"], SS)
        o[f"{net}_lr_depth"] = r.box_average(o[f"{net}_depth"], SS)
    return o
def copies(o):
    for net in ("coarse", "fine"):
        pin[f"{net}_rgb_ori"].copy_(o[f"{net}_comp_rgbs"], non_blocking=True)
        pin[f"{net}_depth_ori"].copy_(o[f"{net}_depth"].view(-1, 1), non_blocking=True)
        pin[f"{net}_rgb"].copy_(o[f"{net}_lr_rgb"], non_blocking=True)
        pin[f"{net}_depth"].copy_(o[f"{net}_lr_depth"].view(-1, 1), non_blocking=True)

def timeit(fn, k=8):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
        torch.cuda.current_stream().synchronize()
    return (time.perf_counter() - t0) / k * 1e3

for name, frame in (("render_frame", frame_new), ("forward_rays+box", frame_old), ("render_frame", frame_new)):
    o = frame(); torch.cuda.synchronize()
    t_frame = timeit(frame)
    t_copy = timeit(lambda: copies(o))
    t_h2d = timeit(lambda: rays_stage.copy_(rays_pinned, non_blocking=True))
    def full():
        rays_stage.copy_(rays_pinned, non_blocking=True)
        copies(frame())
    t_full = timeit(full)
    print(f"{name:18s} frame {t_frame:7.2f} ms | d2h copies {t_copy:6.2f} ms | h2d {t_h2d:5.2f} ms | full step {t_full:7.2f} ms", flush=True)
