"""Device-resident frame rate (160 000 rays) of the option sets the tensor-core path covers beyond the headline 64 + 64 recipe
(dev tool): sample counts, narrow nets, and the fp32 CUDA-core path for comparison."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerf_sr_b200 import Renderer, synthetic as S

rays = S.synthetic_rays(160000, 7, "blender").cuda()
for label, kw, prec in (("64 + 64 (headline)", {}, "bf16x3"), ("64 + 128 (classic NeRF recipe)", dict(N_importance=128), "bf16x3"),
                        ("64 + 192", dict(N_importance=192), "bf16x3"), ("128 + 128", dict(N_coarse=128, N_importance=128), "bf16x3"),
                        ("128 + 0", dict(N_coarse=128, N_importance=0), "bf16x3"), ("64 + 64, W = 128", dict(W=128), "bf16x3"),
                        ("64 + 128, fp32 CUDA-core path", dict(N_importance=128), "fp32_simt")):
    cfg = S.RenderConfig(white_bkgd=True, **kw)
    r = Renderer(cfg, torch.device("cuda:0"), precision=prec)
    r.load_state_dict(0, S.make_mlp_params(cfg, 4)); r.load_state_dict(1, S.make_mlp_params(cfg, 17))
    for _ in range(2):
        r.forward_rays(rays, want_weights=False)
    l0 = r.launch_count
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r.forward_rays(rays, want_weights=False); r.forward_rays(rays, want_weights=False); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 2
    pts = cfg.N_coarse + (cfg.N_coarse + cfg.N_importance if cfg.N_importance else 0)
    print(f"{label:34s} {ms:8.1f} ms/frame = {160000 / ms / 1e3:6.3f} M rays/s = {160000 * pts / ms / 1e3:6.1f} M points/s, {(r.launch_count - l0) // 2} launches", flush=True)
    r.close()
