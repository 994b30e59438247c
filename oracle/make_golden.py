"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Run:  python oracle/make_golden.py

For every fixture it (1) builds synthetic rays + PCG64 kaiming weights
(oracle.make_mlp_params), (2) runs the reference's own forward_rays through
oracle/ref_shim.py, (3) asserts the oracle restatement reproduces the reference
BIT-EXACTLY on CPU (this is what pins the oracle), (4) records the fp64-vs-fp32
noise floor of the algorithm on the same inputs (parity protocol iii), and
(5) stores inputs, rng draws and reference outputs.  Weights are not stored:
tests regenerate them from (seed, cfg) with the same numpy generator.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nerf_oracle as O      # noqa: E402
from oracle import ref_shim              # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> spec.  'args' are reference command-line flags; 'cfg' the matching
# oracle RenderConfig overrides.
FIXTURES = {
    # BASELINE.json configs[0]: single chunk, 64 coarse, no SS, coarse only
    "c1_coarse_only": dict(n=512, rays="blender", seeds=(4, 17),
                           args=["--N_importance", "0", "--white_bkgd"],
                           cfg=dict(N_importance=0, white_bkgd=True)),
    "eval_blender": dict(n=384, rays="blender", seeds=(4, 17),
                         args=["--white_bkgd"], cfg=dict(white_bkgd=True)),
    "eval_llff": dict(n=384, rays="llff", seeds=(21, 8), args=[], cfg=dict()),
    "eval_trained_like": dict(n=256, rays="blender", seeds=(31, 34), bias_std=0.05, sigma_bias=-0.5,
                              args=["--white_bkgd"], cfg=dict(white_bkgd=True)),
    "train_llff_noise": dict(n=256, rays="llff", seeds=(21, 8), train=True, rng_seed=1234,
                             args=["--noise_std", "1.0"], cfg=dict(noise_std=1.0)),
    "train_blender": dict(n=256, rays="blender", seeds=(4, 17), train=True, rng_seed=77,
                          args=["--white_bkgd"], cfg=dict(white_bkgd=True)),
    "opt_lindisp_softplus": dict(n=128, rays="blender", seeds=(4, 17),
                                 args=["--lindisp", "--sigma_activation", "softplus"],
                                 cfg=dict(lindisp=True, sigma_activation="softplus")),
    "opt_color_none": dict(n=128, rays="blender", seeds=(4, 17),
                           args=["--color_activation", "none"],
                           cfg=dict(color_activation="none")),
    "opt_gamma": dict(n=128, rays="llff", seeds=(21, 8),
                      args=["--gamma_correct"], cfg=dict(gamma_correct=True)),
    "opt_nodir": dict(n=128, rays="blender", seeds=(4, 17),
                      args=["--no_dir"], cfg=dict(no_dir=True)),
    "opt_small_net": dict(n=128, rays="blender", seeds=(41, 42),
                          args=["--D", "4", "--W", "128", "--skips", "2", "--deg_pos", "6",
                                "--N_coarse", "32", "--N_importance", "32"],
                          cfg=dict(D=4, W=128, skips=(2,), deg_pos=6, N_coarse=32, N_importance=32)),
    "vanilla_model": dict(n=128, rays="blender", seeds=(4, 17), model="nerf",
                          args=["--white_bkgd"], cfg=dict(white_bkgd=True, viewdir_offset=8)),
}


def _viol(a, b):
    return O.tolerance_violations(a, b)


def build_fixture(name: str, spec: dict) -> dict:
    cfg = O.RenderConfig(**spec["cfg"])
    model_name = spec.get("model", "nerf_downX")
    model, opt = ref_shim.load_reference_model(model_name, spec["args"])
    pc = O.make_mlp_params(cfg, spec["seeds"][0], spec.get("sigma_bias", 0.0), spec.get("bias_std", 0.0))
    pf = O.make_mlp_params(cfg, spec["seeds"][1], spec.get("sigma_bias", 0.0), spec.get("bias_std", 0.0))
    ref_shim.set_weights(model, pc, pf)
    rays = O.synthetic_rays(spec["n"], seed=1000 + spec["seeds"][0], kind=spec["rays"])
    if model_name == "nerf":   # vanilla model reads viewdir from rays[:, 8:11] (models/nerf_model.py:213)
        vd = rays[:, 3:6] / torch.norm(rays[:, 3:6], dim=-1, keepdim=True)
        rays = torch.cat([rays, vd], 1).contiguous()

    rng = None
    with torch.no_grad():
        if spec.get("train"):
            model.randomized = True
            torch.manual_seed(spec["rng_seed"])
            ref = model.forward_rays(rays)
            g = torch.Generator().manual_seed(spec["rng_seed"])
            rng = O.RenderRng.draw(rays.shape[0], cfg, g)
        else:
            ref = model.forward_rays(rays)
        extras: dict = {}
        ora = O.forward_rays(pc, pf, rays, cfg, rng, extras=extras)
    for k in ref:   # the pin: restatement == reference, bit for bit
        assert torch.equal(ref[k], ora[k]), f"{name}: oracle != reference on {k}: " \
            f"{float((ref[k]-ora[k]).abs().max())}"

    # fp64 noise floor of the algorithm itself (protocol iii)
    with torch.no_grad():
        pc64 = {k: v.double() for k, v in pc.items()}
        pf64 = {k: v.double() for k, v in pf.items()}
        rng64 = None
        if rng is not None:
            rng64 = O.RenderRng(*[None if t is None else t.double() for t in
                                  (rng.u_coarse, rng.noise_coarse, rng.u_fine, rng.noise_fine)])
        o64 = O.forward_rays(pc64, pf64, rays.double(), cfg, rng64)
    floor = {k: _viol(ref[k], o64[k].float()) for k in ref}

    meta = dict(name=name, cfg=spec["cfg"], seeds=list(spec["seeds"]),
                sigma_bias=spec.get("sigma_bias", 0.0), bias_std=spec.get("bias_std", 0.0),
                model=model_name, train=bool(spec.get("train")), torch=torch.__version__,
                numpy=np.__version__, reference_args=spec["args"],
                fp64_floor={k: dict(max_abs=v[0], viol=v[1]) for k, v in floor.items()},
                mean_opacity={k: float(ref[k].mean()) for k in ref if k.endswith("opacity")})
    arrays = {"rays": rays.numpy()}
    for k, v in ref.items():
        arrays["out_" + k] = v.numpy()
    arrays["z_coarse"] = extras["z_coarse"].numpy()
    arrays["raw_coarse"] = extras["raw_coarse"].numpy()
    if "z_fine" in extras:
        arrays["z_fine"] = extras["z_fine"].numpy()
        arrays["raw_fine"] = extras["raw_fine"].numpy()
    if rng is not None:
        for f in ("u_coarse", "noise_coarse", "u_fine", "noise_fine"):
            t = getattr(rng, f)
            if t is not None:
                arrays["rng_" + f] = t.numpy()
    arrays["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return arrays


def build_raygen_fixture() -> dict:
    """a1-a4 against the reference's own helpers (models/utils.py) and einops
    grouping string (data/blender_downX_dataset.py:213-215)."""
    ref_shim._install_stubs()
    if ref_shim.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    from models.utils import get_ray_directions, get_rays, get_ndc_rays
    from einops import rearrange
    out = {}
    rng = np.random.default_rng(7)
    for tag, (H, W, s, focal, ndc) in {"blender": (24, 32, 2, 41.7, False),
                                        "llff": (36, 24, 4, 29.3, True)}.items():
        if ndc:   # forward-facing (LLFF) pose: small rotation about identity so d_z stays away from 0
            a = np.eye(3) + 0.15 * rng.standard_normal((3, 3))
        else:
            a = rng.standard_normal((3, 3))
        q, _ = np.linalg.qr(a)
        q = q * np.sign(np.diag(q))[None, :] if ndc else q
        c2w = np.concatenate([q, rng.standard_normal((3, 1)) * 0.5 + np.array([[0.], [0.], [2.5]])], 1)
        c2w = torch.from_numpy(c2w.astype(np.float32))
        dirs = get_ray_directions(H, W, focal)
        o, d = get_rays(dirs, c2w)
        if ndc:
            o, d = get_ndc_rays(H, W, focal, 1.0, o, d)
            near, far = torch.zeros_like(o[:, :1]), torch.ones_like(o[:, :1])
        else:
            near, far = 2.0 * torch.ones_like(o[:, :1]), 6.0 * torch.ones_like(o[:, :1])
        rays = torch.cat([o, d, near, far], 1).view(H, W, 8)
        rays = rearrange(rays, "(h s1) (w s2) c -> (h w) (s1 s2) c", s1=s, s2=s).reshape(-1, 8)
        mine = O.build_frame_rays(c2w, H, W, focal, s, 2.0, 6.0, ndc)
        assert torch.equal(rays, mine), tag
        out[f"{tag}_c2w"] = c2w.numpy()
        out[f"{tag}_params"] = np.array([H, W, s, focal, float(ndc), 2.0, 6.0], np.float64)
        out[f"{tag}_rays"] = rays.numpy()
    return out


def main() -> None:
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, spec in FIXTURES.items():
        arrays = build_fixture(name, spec)
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **arrays)
        meta = json.loads(bytes(arrays["meta_json"]).decode())
        print(f"{name:24s} {os.path.getsize(path)/1e3:8.1f} KB  opacity={meta['mean_opacity']}  "
              f"fp64 floor fine_rgb={meta['fp64_floor'].get('fine_comp_rgbs')}")
    rg = build_raygen_fixture()
    np.savez_compressed(os.path.join(GOLDEN, "raygen.npz"), **rg)
    print("raygen ok")


if __name__ == "__main__":
    main()
