"""Dev tool: event-timed backward of one 2048-ray batch under the chain's debug flags (8 no dZ stores, 16 no mask loads,
32 no weight stream), with the chain's in-kernel SM clock -> cycles per tile.  Timing only: flags give WRONG gradients."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_sr_b200 import Renderer, Trainer
from nerf_sr_b200.synthetic import RenderConfig, make_mlp_params, synthetic_rays
dev = torch.device("cuda:0")
cfg = RenderConfig(noise_std=1.0)
r = Renderer(cfg, dev, precision="bf16x3")
tr = Trainer(r, make_mlp_params(cfg, 21), make_mlp_params(cfg, 8), downscale=2)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rays = synthetic_rays(n, 5, "llff").to(dev)
tgt = torch.rand(n // 4, 3, device=dev)
gen = torch.Generator(device=dev).manual_seed(0)
rng = tr.draw_rng(n, gen)
out = r.render_train(rays, rng, want_weights=False)
_, _, g_c = r.lr_loss_grad(out["coarse_comp_rgbs"], tgt, 2, 1.0)
_, _, g_f = r.lr_loss_grad(out["fine_comp_rgbs"], tgt, 2, 1.0)
grads = {"coarse_comp_rgbs": g_c, "fine_comp_rgbs": g_f}
for flags in (0, 8, 32, 40, 16):
    r.lib.nsr_debug_set_flags(r._h, flags)
    for _ in range(3):
        r.backward(rays, rng, grads)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        r.backward(rays, rng, grads)
    b.record()
    torch.cuda.synchronize()
    print(json.dumps({"flags": flags, "backward_ms": round(a.elapsed_time(b) / 10, 4), "chain_fine_sm_mhz": round(r.kernel_clock_mhz() or 0, 1)}))
r.lib.nsr_debug_set_flags(r._h, 0)
