"""Time the fused training iteration (BASELINE config 3 shape: 512 LR pixels x 2x2 = 2048 rays per step,
64 + 128 samples, LLFF-like rays, sigma noise) phase by phase with CUDA events.  Dev tool; prints JSON."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_sr_b200 import Renderer, Trainer                      # noqa: E402
from nerf_sr_b200.synthetic import RenderConfig, make_mlp_params, synthetic_rays   # noqa: E402


def main():
    n_lr = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    prec = sys.argv[3] if len(sys.argv) > 3 else "bf16x3"
    dev = torch.device("cuda:0")
    cfg = RenderConfig(noise_std=1.0)
    pc, pf = make_mlp_params(cfg, 21), make_mlp_params(cfg, 8)
    r = Renderer(cfg, dev, precision=prec)
    r.lib.nsr_debug_set_flags(r._h, int(os.environ.get("NSR_DEBUG_FLAGS", "0")))    # e.g. 4: dW GEMMs on one stream
    tr = Trainer(r, pc, pf, downscale=2)
    rays = synthetic_rays(n_lr * 4, 5, "llff").to(dev)
    tgt = torch.rand(n_lr, 3, device=dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        tr.optimize_parameters(rays, tgt, tr.draw_rng(rays.shape[0], gen))
    torch.cuda.synchronize()
    l0 = r.launch_count
    t = {"rng": 0.0, "forward": 0.0, "loss": 0.0, "backward": 0.0, "adam_pack": 0.0}
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(steps):
        tr.optimize_parameters(rays, tgt, tr.draw_rng(rays.shape[0], gen))
    e1.record()
    torch.cuda.synchronize()
    whole = e0.elapsed_time(e1) / steps
    launches = (r.launch_count - l0) / steps
    for _ in range(steps):
        a, b, c, d, e, f = ev(), ev(), ev(), ev(), ev(), ev()
        a.record()
        rng = tr.draw_rng(rays.shape[0], gen)
        b.record()
        out = r.render_train(rays, rng, want_weights=False)
        c.record()
        _, mc, g_c = r.lr_loss_grad(out["coarse_comp_rgbs"], tgt, 2, 1.0)
        _, mf, g_f = r.lr_loss_grad(out["fine_comp_rgbs"], tgt, 2, 1.0)
        d.record()
        gc, gf = r.backward(rays, rng, {"coarse_comp_rgbs": g_c, "fine_comp_rgbs": g_f})
        e.record()
        tr.step += 1
        for w, g in ((0, gc), (1, gf)):
            r.adam_step(tr.params[w], g, tr.m[w], tr.v[w], tr.step, tr.lr)
            r.load_params(w, tr.params[w])
        f.record()
        torch.cuda.synchronize()
        for k, (x, y) in zip(t, ((a, b), (b, c), (c, d), (d, e), (e, f))):
            t[k] += x.elapsed_time(y) / steps
    flop = rays.shape[0] * 227.87e6 * 3            # forward + 2x backward
    print(json.dumps({"rays_per_step": rays.shape[0], "precision": prec, "ms_per_step": whole,
                      "train_rays_per_s": rays.shape[0] / whole * 1e3, "algorithmic_tflops": flop / whole / 1e9,
                      "launches_per_step": launches, "phase_ms": t,
                      "train_ws_gb": r.lib.nsr_train_workspace_bytes(r._h, rays.shape[0]) / 1e9}))


if __name__ == "__main__":
    main()
