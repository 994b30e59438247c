"""Generate tests/golden/train_step_*.npz from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Run:  python oracle/make_golden_train.py

Pins oracle/train_oracle.py (scope row f-1: loss, backward, gradient clipping, Adam) against the
reference's own ``optimize_parameters`` (models/nerf_downX_model.py:398-408) run for TWO consecutive
iterations on CPU: losses, every gradient tensor and every updated parameter must be BIT-EQUAL.
The fixture then stores the inputs, the RNG draws, the losses and -- to stay small -- a fixed
pseudo-random subset of 256 entries plus the 2-norm of every gradient / updated-parameter tensor
(tests recompute the full tensors with the oracle and check them against these).
It also records the fp64-vs-fp32 noise floor of the gradients (relative L2 per tensor)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nerf_oracle as O      # noqa: E402
from oracle import train_oracle as T     # noqa: E402
from oracle import ref_shim              # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
N_SUB = 256

FIXTURES = {
    "train_step_blender": dict(n_lr=64, s=2, rays="blender", seeds=(4, 17), rng_seed=2024,
                               args=["--white_bkgd"], cfg=dict(white_bkgd=True), tcfg=dict()),
    "train_step_llff_clip": dict(n_lr=48, s=2, rays="llff", seeds=(21, 8), rng_seed=99,
                                 args=["--noise_std", "1.0", "--grad_clip_val", "0.05", "--lr", "1e-3",
                                       "--lambda_coarse_mse", "0.5"],
                                 cfg=dict(noise_std=1.0),
                                 tcfg=dict(grad_clip_val=0.05, lr=1e-3, lambda_coarse_mse=0.5)),
    "train_step_s4_value_clip": dict(n_lr=12, s=4, rays="blender", seeds=(31, 34), rng_seed=5,
                                     args=["--white_bkgd", "--downscale", "4", "--grad_clip_val", "1e-4",
                                           "--grad_clip_type", "value"],
                                     cfg=dict(white_bkgd=True, downscale=4),
                                     tcfg=dict(grad_clip_val=1e-4, grad_clip_type="value")),
    # sub-pixel variance terms (--use_var_loss, --use_depth_var_loss; models/nerf_downX_model.py:331-335,349-353,374-378)
    "train_step_var_losses": dict(n_lr=48, s=2, rays="blender", seeds=(4, 17), rng_seed=77,
                                  args=["--white_bkgd", "--use_var_loss", "--lambda_coarse_var", "0.02",
                                        "--use_depth_var_loss", "--lambda_fine_depth_var", "0.05"],
                                  cfg=dict(white_bkgd=True),
                                  tcfg=dict(use_var_loss=True, lambda_coarse_var=0.02, use_depth_var_loss=True,
                                            lambda_fine_depth_var=0.05)),
    # all loss terms of the fused epilogue at once, 4x4 SS, LLFF-like rays with sigma noise, SISR target (--sisr_path, :364-367)
    "train_step_sr_var_s4": dict(n_lr=10, s=4, rays="llff", seeds=(21, 8), rng_seed=11, sisr=True,
                                 args=["--noise_std", "1.0", "--downscale", "4", "--use_var_loss", "--use_depth_var_loss",
                                       "--sisr_path", "/nonexistent/sisr", "--grad_clip_val", "0.1"],
                                 cfg=dict(noise_std=1.0, downscale=4),
                                 tcfg=dict(use_var_loss=True, use_depth_var_loss=True, grad_clip_val=0.1)),
    # reference-view term (--with_ref, :321-324,369-372): a second forward over 40 sub-pixel rays of the reference image
    "train_step_with_ref": dict(n_lr=32, s=2, rays="llff", seeds=(21, 8), rng_seed=31, n_ref=40,
                                args=["--noise_std", "1.0", "--with_ref", "--use_var_loss"],
                                cfg=dict(noise_std=1.0), tcfg=dict(use_var_loss=True)),
}


def summarize(tensors, prefix: str, arrays: dict) -> None:
    for i, t in enumerate(tensors):
        flat = t.detach().reshape(-1)
        idx = T.golden_sub_indices(flat.numel(), i, N_SUB)
        arrays[f"{prefix}_{i}_sub"] = flat[torch.from_numpy(idx)].numpy()
        arrays[f"{prefix}_{i}_norm"] = np.array([float(torch.linalg.vector_norm(flat.double()))])


def build(name: str, spec: dict) -> dict:
    cfg = O.RenderConfig(**spec["cfg"])
    tcfg = T.TrainConfig(**spec["tcfg"])
    s = spec["s"]
    n = spec["n_lr"] * s * s
    model, opt = ref_shim.load_reference_model("nerf_downX", spec["args"], train=True)
    pc = O.make_mlp_params(cfg, spec["seeds"][0])
    pf = O.make_mlp_params(cfg, spec["seeds"][1])
    ref_shim.set_weights(model, pc, pf)
    if tcfg.use_depth_var_loss:
        # models/nerf_downX_model.py:351 divides a grad-requiring tensor by ``self.far``, a shape-(1,) numpy array (:284);
        # torch refuses that (Tensor.__array__ on a tensor that requires grad), so the flag cannot run in train mode as
        # written.  The harness -- not the reference tree -- hands the model the same value as a Python float, the
        # scalar the expression means; everything else is the unmodified reference.
        ref_forward_rays = model.forward_rays

        def forward_rays_scalar_far(r):
            out = ref_forward_rays(r)
            model.far = float(model.far[0])
            return out
        model.forward_rays = forward_rays_scalar_far
    rays = O.synthetic_rays(n, seed=3000 + spec["seeds"][0], kind=spec["rays"])
    tg = torch.Generator().manual_seed(spec["rng_seed"] + 1)
    target = torch.rand(spec["n_lr"], 3, generator=tg)
    target_sr = torch.rand(n, 3, generator=tg) if spec.get("sisr") else None
    n_ref = spec.get("n_ref", 0)
    ref_rays = O.synthetic_rays(n_ref, seed=4000 + spec["seeds"][0], kind=spec["rays"]) if n_ref else None
    ref_rgbs = torch.rand(n_ref, 3, generator=tg) if n_ref else None

    state = T.TrainState(pc, pf)
    g = torch.Generator().manual_seed(spec["rng_seed"])
    torch.manual_seed(spec["rng_seed"])
    arrays = {"rays": rays.numpy(), "target": target.numpy()}
    if target_sr is not None:
        arrays["target_sr"] = target_sr.numpy()
    if n_ref:
        arrays["ref_rays"], arrays["ref_rgbs"] = ref_rays.numpy(), ref_rgbs.numpy()
    meta = dict(name=name, cfg=spec["cfg"], tcfg=spec["tcfg"], seeds=list(spec["seeds"]), s=s,
                reference_args=spec["args"], torch=torch.__version__, steps=[])
    ref_params = lambda: [p for p in model.netCoarse.parameters()] + [p for p in model.netFine.parameters()]
    for step in range(2):
        # ---- reference iteration ----
        model.set_input({"rays": rays.clone(), "rgbs": target.clone(),
                         **({"rgbs_sr": target_sr.clone()} if target_sr is not None else {}),
                         **({"ref_rays": ref_rays.clone(), "ref_rgbs": ref_rgbs.clone()} if n_ref else {})})
        model.optimize_parameters()
        ref_grads = [p.grad.detach().clone() for p in ref_params()]
        ref_losses = dict(coarse_mse=model.loss_coarse_mse.detach(), fine_mse=model.loss_fine_mse.detach(),
                          tot=model.loss_tot.detach(), coarse_psnr=model.loss_coarse_psnr.detach(),
                          fine_psnr=model.loss_fine_psnr.detach())
        if tcfg.use_var_loss:
            ref_losses.update(coarse_var=model.loss_out_coarse_var.detach(), fine_var=model.loss_out_fine_var.detach())
        if tcfg.use_depth_var_loss:
            ref_losses.update(coarse_depth_var=model.loss_coarse_depth_var.detach(),
                              fine_depth_var=model.loss_fine_depth_var.detach())
        if target_sr is not None:
            ref_losses.update(coarse_mse_sr=model.loss_coarse_mse_sr.detach(), fine_mse_sr=model.loss_fine_mse_sr.detach())
        if n_ref:
            ref_losses.update(ref_coarse_mse=model.loss_ref_coarse_mse.detach(), ref_fine_mse=model.loss_ref_fine_mse.detach())
        # ---- oracle iteration on the same draws ----
        rng = O.RenderRng.draw(n, cfg, g)
        ref_rng = O.RenderRng.draw(n_ref, cfg, g) if n_ref else None
        if step == 0:   # fp64 floor of the gradients at the initial weights
            refkw = dict(ref_rays=ref_rays, ref_rgbs=ref_rgbs, ref_rng=ref_rng) if n_ref else {}
            _, gc32, gf32, _ = T.loss_and_grads(state.pc, state.pf, rays, target, cfg, tcfg, rng, s, target_sr=target_sr, **refkw)
            d = lambda t: None if t is None else t.double()
            rng64 = O.RenderRng(d(rng.u_coarse), d(rng.noise_coarse), d(rng.u_fine), d(rng.noise_fine))
            refkw64 = dict(ref_rays=ref_rays.double(), ref_rgbs=ref_rgbs.double(),
                           ref_rng=O.RenderRng(d(ref_rng.u_coarse), d(ref_rng.noise_coarse), d(ref_rng.u_fine),
                                               d(ref_rng.noise_fine))) if n_ref else {}
            _, gc64, gf64, _ = T.loss_and_grads({k: v.double() for k, v in state.pc.items()},
                                                {k: v.double() for k, v in state.pf.items()},
                                                rays.double(), target.double(), cfg, tcfg, rng64, s,
                                                target_sr=None if target_sr is None else target_sr.double(), **refkw64)
            rel = lambda a, b: float(torch.linalg.vector_norm(a.double() - b) / (torch.linalg.vector_norm(b) + 1e-30))
            meta["fp64_floor_rel_l2"] = dict(
                coarse={k: rel(gc32[k], gc64[k]) for k in gc32}, fine={k: rel(gf32[k], gf64[k]) for k in gf32})
        losses, grads = T.optimize_parameters(state, rays, target, cfg, tcfg, rng, s, target_sr=target_sr,
                                              ref_rays=ref_rays, ref_rgbs=ref_rgbs, ref_rng=ref_rng)
        # ---- the pin ----
        for k in ref_losses:
            assert torch.equal(ref_losses[k], losses[k]), (name, step, k, float(ref_losses[k]), float(losses[k]))
        for i, (a, b) in enumerate(zip(ref_grads, grads)):
            assert torch.equal(a, b), (name, step, "grad", i, float((a - b).abs().max()))
        for i, (a, b) in enumerate(zip(ref_params(), state.param_list())):
            assert torch.equal(a.detach(), b), (name, step, "param", i, float((a.detach() - b).abs().max()))
        for f in ("u_coarse", "noise_coarse", "u_fine", "noise_fine"):
            t = getattr(rng, f)
            if t is not None:
                arrays[f"rng{step}_{f}"] = t.numpy()
            t = getattr(ref_rng, f) if n_ref else None
            if t is not None:
                arrays[f"refrng{step}_{f}"] = t.numpy()
        summarize(grads, f"grad{step}", arrays)
        summarize(state.param_list(), f"param{step}", arrays)
        meta["steps"].append({k: float(v) for k, v in losses.items()})
    # scheduler rule (models/networks.py:102-113) against the reference's LambdaLR
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    from models.networks import get_scheduler
    opt.lr_policy, opt.lr_final, opt.n_epochs, opt.n_epochs_decay = tcfg.lr_policy, tcfg.lr_final, 3, 4
    tc2 = T.TrainConfig(**{**spec["tcfg"], "n_epochs": 3, "n_epochs_decay": 4})
    sched = get_scheduler(model.optimizer, opt, -1)
    lrs = []
    for e in range(8):
        lrs.append(model.optimizer.param_groups[0]["lr"])
        assert abs(lrs[-1] - T.lr_at_epoch(tc2, e)) <= 1e-12 * tc2.lr, (e, lrs[-1], T.lr_at_epoch(tc2, e))
        model.optimizer.step()
        sched.step()
    meta["lr_schedule_n3_d4"] = lrs
    arrays["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return arrays


def main() -> None:
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    only = sys.argv[1:]
    for name, spec in FIXTURES.items():
        if only and name not in only:
            continue
        arrays = build(name, spec)
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **arrays)
        meta = json.loads(bytes(arrays["meta_json"]).decode())
        fl = meta["fp64_floor_rel_l2"]
        print(f"{name:28s} {os.path.getsize(path)/1e3:8.1f} KB  steps={meta['steps']}  "
              f"fp64 floor max rel-L2: coarse {max(fl['coarse'].values()):.2e} fine {max(fl['fine'].values()):.2e}")


if __name__ == "__main__":
    main()
