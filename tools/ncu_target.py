"""Short single-GPU target for ncu captures: two full frames through the one-launch kernel (first = warm-up), then the same
frame as separate coarse / fine launches (debug flag 64)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from nerf_sr_b200 import Renderer  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
cfg, pc, pf, rays = bench.make_inputs()
r = Renderer(cfg, torch.device("cuda:0"), precision=prec)
r.load_state_dict(0, pc)
r.load_state_dict(1, pf)
rays = rays.cuda()
for _ in range(2):
    r.render_frame(rays, bench.SS)
r.set_debug_flags(64)
r.forward_rays(rays, want_weights=False)
torch.cuda.synchronize()
