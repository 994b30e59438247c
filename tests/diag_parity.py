"""GPU diagnostic: detailed per-stage error statistics of the CUDA path against the golden
fixtures (max abs error, tolerance-violation fraction).  Not a test: prints, never asserts, so
one gpurun call shows everything.  Usage: python tests/diag_parity.py [precision ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from conftest import Fixture, golden_names  # noqa: E402
from nerf_sr_b200 import NsrError, Renderer  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402


def stat(a, b):
    mx, v = O.tolerance_violations(a.cpu(), b)
    return f"{mx:.2e}/{100*v:.2f}%"


def main():
    precs = sys.argv[1:] or ["fp32_simt", "bf16x3", "fp16x3"]
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0), flush=True)
    for prec in precs:
        for name in golden_names():
            fx = Fixture(name)
            try:
                r = Renderer(fx.cfg, dev, precision=prec, viewdir_offset=fx.cfg.viewdir_offset)
            except NsrError as e:
                print(f"[{prec}] {name}: skipped ({e})", flush=True)
                continue
            r.load_state_dict(0, fx.p_coarse)
            r.load_state_dict(1, fx.p_fine)
            rays = fx.rays.to(dev)
            rng = None
            if fx.rng is not None:
                rng = {k: getattr(fx.rng, k) for k in ("u_coarse", "noise_coarse", "u_fine", "noise_fine")
                       if getattr(fx.rng, k) is not None}
            t0 = time.time()
            out = r.forward_rays(rays, rng, want_z_fine=True)
            torch.cuda.synchronize()
            dt = time.time() - t0
            line = [f"[{prec}] {name} ({dt*1e3:.0f} ms)"]
            for k in ("coarse_comp_rgbs", "coarse_depth", "coarse_opacity", "coarse_weights"):
                line.append(f"{k[7:]}={stat(out[k], fx.out[k])}")
            # raw MLP output of the coarse net
            nz = fx.rng.noise_coarse.to(dev) if (fx.rng is not None and fx.rng.noise_coarse is not None) else None
            pc = r.render_pass(0, rays, fx.z_coarse.to(dev), nz, want_raw=True)
            line.append(f"raw_c={stat(pc['raw'], fx.raw_coarse)}")
            if fx.cfg.N_importance > 0:
                line.append("| e2e fine:")
                for k in ("fine_comp_rgbs", "fine_depth", "fine_weights"):
                    line.append(f"{k[5:]}={stat(out[k], fx.out[k])}(floor {100*fx.meta['fp64_floor'][k]['viol']:.2f}%)")
                line.append(f"z_fine={stat(out['z_fine'], fx.z_fine)}")
                # resampler unit parity and teacher-forced fine pass
                uf = fx.rng.u_fine.to(dev) if (fx.rng is not None and fx.rng.u_fine is not None) else None
                zf = r.resample_along_rays(fx.z_coarse.to(dev), fx.out["coarse_weights"].to(dev), uf)
                line.append(f"| resample={stat(zf, fx.z_fine)}")
                nzf = fx.rng.noise_fine.to(dev) if (fx.rng is not None and fx.rng.noise_fine is not None) else None
                pf = r.render_pass(1, rays, fx.z_fine.to(dev), nzf, want_raw=True)
                line.append("| teacher-forced fine:")
                for kl, kr in (("comp_rgbs", "fine_comp_rgbs"), ("depth", "fine_depth"), ("opacity", "fine_opacity"),
                               ("weights", "fine_weights")):
                    line.append(f"{kl}={stat(pf[kl], fx.out[kr])}")
                line.append(f"raw_f={stat(pf['raw'], fx.raw_fine)}")
            print(" ".join(line), flush=True)
            r.close()


if __name__ == "__main__":
    main()
