"""Dev diagnostic: one forward + backward of a committed training fixture (sync after every stage), for
running under compute-sanitizer / CUDA_LAUNCH_BLOCKING on the GPU box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import TrainFixture, rng_dict          # noqa: E402
from nerf_sr_b200 import Renderer, Trainer           # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "train_step_blender"
    fx = TrainFixture(name)
    dev = torch.device("cuda:0")
    r = Renderer(fx.cfg, dev, precision="bf16x3")
    tr = Trainer(r, fx.p_coarse, fx.p_fine, downscale=fx.s)
    rng = rng_dict(fx.rng[0])
    rays, tgt = fx.rays.to(dev), fx.target.to(dev)
    out = r.render_train(rays, rng, want_weights=False)
    torch.cuda.synchronize(); print("forward ok", flush=True)
    _, mc, g_c = r.lr_loss_grad(out["coarse_comp_rgbs"], tgt, fx.s, 1.0)
    _, mf, g_f = r.lr_loss_grad(out["fine_comp_rgbs"], tgt, fx.s, 1.0)
    torch.cuda.synchronize(); print("loss ok", mc.tolist(), mf.tolist(), flush=True)
    gc, gf = r.backward(rays, rng, {"coarse_comp_rgbs": g_c, "fine_comp_rgbs": None})
    torch.cuda.synchronize(); print("backward coarse ok", float(gc.norm()), flush=True)
    gc, gf = r.backward(rays, rng, {"coarse_comp_rgbs": g_c, "fine_comp_rgbs": g_f})
    torch.cuda.synchronize(); print("backward ok", float(gc.norm()), float(gf.norm()), flush=True)


if __name__ == "__main__":
    main()
