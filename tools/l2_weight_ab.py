"""Dev tool: is the weight stream L2 -> shared memory the limiter of k_tc_pass?  Times the fused forward (no stash) at a
small and at a frame-sized batch with the producer's bulk copies on (flags 0) and skipped after the first ring fill
(debug flag 1: stale weights, WRONG results, timing only), with the SM clock measured inside the kernel."""
import json
import sys
import os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_sr_b200 import Renderer
from nerf_sr_b200.synthetic import RenderConfig, make_mlp_params, synthetic_rays

dev = torch.device("cuda:0")
cfg = RenderConfig(white_bkgd=True)
r = Renderer(cfg, dev, precision="bf16x3")
r.load_state_dict(0, make_mlp_params(cfg, 4))
r.load_state_dict(1, make_mlp_params(cfg, 17))
for n in (2048, 160000):
    rays = synthetic_rays(n, 1, "blender").to(dev)
    for flags in (0, 1):
        r.lib.nsr_debug_set_flags(r._h, flags)
        for _ in range(3):
            r.forward_rays(rays, want_weights=False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k = 20 if n < 10000 else 5
        a.record()
        for _ in range(k):
            r.forward_rays(rays, want_weights=False)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / k
        tiles = n * 192 / 128
        print(json.dumps({"rays": n, "flags": flags, "ms": round(ms, 4), "us_per_tile_per_sm": round(ms * 1e3 / (tiles / 148), 2),
                          "sm_mhz_in_kernel": round(r.kernel_clock_mhz() or 0, 1),
                          "weight_stream_TBps": round(tiles * 72 * 32768 / (ms * 1e-3) / 1e12, 2)}))
r.lib.nsr_debug_set_flags(r._h, 0)
