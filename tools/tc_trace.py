"""Timeline of one tile of the tcgen05 kernel (trace build): NSR_LIB_PATH=.../libnsr_b200_trace.so.
Prints per-warp (tag, delta-cycles) sequences for CTA 0, tile iteration 3 of the fine pass."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NSR_LIB_PATH", os.path.join(ROOT, "nerf_sr_b200", "libnsr_b200_trace.so"))
import torch  # noqa: E402

import bench  # noqa: E402
from nerf_sr_b200 import Renderer  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
cfg, pc, pf, rays = bench.make_inputs()
r = Renderer(cfg, torch.device("cuda:0"), precision=prec)
r.load_state_dict(0, pc)
r.load_state_dict(1, pf)
rays = rays.cuda()
out = r.forward_rays(rays, want_weights=False, want_z_fine=True)
buf = torch.zeros(16 * 2048, dtype=torch.int64, device="cuda")
r.lib.nsr_debug_set_trace(r._h, buf.data_ptr())
r.render_pass(1, rays, out["z_fine"])
torch.cuda.synchronize()
b = buf.cpu().view(16, 1024, 2)
t0 = min(int(b[w, 0, 1]) for w in range(16) if int(b[w, 0, 0]) != 0)
for w in range(16):
    ev = [(int(b[w, i, 0]), int(b[w, i, 1]) - t0) for i in range(1024) if int(b[w, i, 0]) != 0]
    if not ev:
        continue
    print(f"--- warp {w}: {len(ev)} events")
    prev = ev[0][1]
    line = []
    for tag, t in ev:
        line.append(f"{tag}@{t}(+{t - prev})")
        prev = t
    for i in range(0, len(line), 8):
        print("   " + " ".join(line[i:i + 8]))
