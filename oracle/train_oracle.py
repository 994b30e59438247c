"""CPU oracle for the training step around the render hot path (scope row f-1).

TEST INFRASTRUCTURE ONLY (see oracle/nerf_oracle.py header): imported by tests/, smoke() and the
baseline legs of bench.py, never by nerf_sr_b200/.

Restates, with plain torch CPU ops, what the reference does in one ``optimize_parameters`` call
(models/nerf_downX_model.py:398-408):

  forward()                       :316-319  (train mode: stratified jitter, sigma noise)
  comp_low_res_output()           :326-353  (s x s box average of the composite colours)
  calculate_losses()              :355-388  (lambda_c * MSE(coarse_lr, target) + lambda_f * MSE(fine_lr, target),
                                             PSNR of both; + the sub-pixel variance terms :331-335,349-353,
                                             374-378, the SR-target term :364-367 and the reference-view term
                                             :321-324,369-372 (a second forward over ``data_ref_rays``))
  loss_tot.backward()             :390-396  (torch autograd through exactly the ops nerf_oracle restates;
                                             the fine z-values use coarse_weights.detach(), :302)
  clip_grad_norm_/clip_grad_value_:403-407
  optimizer.step()                :408      (torch.optim.Adam, betas=(beta1, 0.999), eps 1e-8, :201-204)
  scheduler (per epoch)           models/networks.py:102-118  ('linear' | 'exp' LambdaLR rules)

Gradients are DEFINED by autograd over ``nerf_oracle.forward_rays`` -- the same definition the
reference uses -- so the pin (oracle/make_golden_train.py) is bit-exact on CPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import nerf_oracle as O

Tensor = torch.Tensor


@dataclass
class TrainConfig:
    """Subset of the reference's train options the step reads (options/train_options.py:33-55,
    models/nerf_model.py lambda_* flags)."""
    lr: float = 5e-4
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    lambda_coarse_mse: float = 1.0
    lambda_fine_mse: float = 1.0
    grad_clip_val: float = 0.0
    grad_clip_type: str = "norm"        # norm | value
    lr_policy: str = "exp"              # linear | exp
    lr_final: float = 5e-6
    n_epochs: int = 20
    n_epochs_decay: int = 10
    # NeRFDownXModel.modify_commandline_options (models/nerf_downX_model.py:107-112)
    use_var_loss: bool = False
    lambda_coarse_var: float = 0.01
    lambda_fine_var: float = 0.01
    use_depth_var_loss: bool = False
    lambda_coarse_depth_var: float = 0.01
    lambda_fine_depth_var: float = 0.01


def loss_and_grads(pc: Dict[str, Tensor], pf: Dict[str, Tensor], rays: Tensor, target_lr: Tensor,
                   cfg: O.RenderConfig, tcfg: TrainConfig, rng: Optional[O.RenderRng] = None,
                   s: int = 2, z_fine_override: Optional[Tensor] = None, extras: Optional[dict] = None,
                   target_sr: Optional[Tensor] = None, ref_rays: Optional[Tensor] = None, ref_rgbs: Optional[Tensor] = None,
                   ref_rng: Optional[O.RenderRng] = None, ref_z_fine_override: Optional[Tensor] = None,
                   relu_masks=None):
    """One forward + backward.  Returns (losses dict, grads_coarse dict, grads_fine dict, outputs dict).
    target_lr: [N/s^2, 3].  target_sr: [N, 3] or None = ``data_rgbs_sr`` (``--sisr_path``, :364-367).
    ref_rays [M, 8] / ref_rgbs [M, 3] (``--with_ref``, :321-324, :369-372): sub-pixel rays of the reference view with
    their HR colours; rendered by a second forward after the main one (so its train-mode draws come second).
    Gradients are None-free: parameters the loss does not reach get zeros
    (the reference leaves .grad = None for them; Adam then skips the parameter)."""
    pc_r = {k: v.detach().clone().requires_grad_(True) for k, v in pc.items()}
    pf_r = {k: v.detach().clone().requires_grad_(True) for k, v in pf.items()}
    out = O.forward_rays(pc_r, pf_r, rays, cfg, rng, z_fine_override=z_fine_override, extras=extras, relu_masks=relu_masks)
    out_ref = None
    if ref_rays is not None:                                                  # :321-324
        out_ref = O.forward_rays(pc_r, pf_r, ref_rays, cfg, ref_rng, z_fine_override=ref_z_fine_override)
    # comp_low_res_output (:326-353).  The ops are created in the reference's order: autograd sums the branches
    # that meet at one output in reverse creation order, and fp32 addition of three terms is not associative.
    n_lr = target_lr.shape[0]
    fine = cfg.N_importance > 0
    grp = lambda x: torch.reshape(x, (n_lr, s * s, -1))
    losses = {}
    c_rgb_ori = out["coarse_comp_rgbs"].clone()                               # :327
    if tcfg.use_var_loss:                                                     # :331-335
        losses["coarse_var"] = subpixel_variance_sum(out["coarse_comp_rgbs"], n_lr, s)
        losses["fine_var"] = subpixel_variance_sum(out["fine_comp_rgbs"], n_lr, s)
    lr_c = torch.mean(grp(out["coarse_comp_rgbs"]), dim=1)                    # :337-338
    c_depth_ori = out["coarse_depth"].clone()                                 # :339
    if fine:
        f_rgb_ori = out["fine_comp_rgbs"].clone()                             # :343
        lr_f = torch.mean(grp(out["fine_comp_rgbs"]), dim=1)                  # :344-345
        f_depth_ori = out["fine_depth"].clone()                               # :346
    if tcfg.use_depth_var_loss:                                               # :349-353
        far = float(rays[0, 7])                                               # self.far, :284 (see subpixel_variance_sum)
        losses["coarse_depth_var"] = subpixel_variance_sum(c_depth_ori, n_lr, s, far)
        losses["fine_depth_var"] = subpixel_variance_sum(f_depth_ori, n_lr, s, far)
    # calculate_losses (:355-378)
    loss_c = torch.nn.functional.mse_loss(lr_c, target_lr) * tcfg.lambda_coarse_mse   # :357
    losses["coarse_mse"] = loss_c
    tot = loss_c
    if fine:
        loss_f = torch.nn.functional.mse_loss(lr_f, target_lr) * tcfg.lambda_fine_mse  # :359
        losses["fine_mse"] = loss_f
        tot = tot + loss_f                                                    # :362
    if target_sr is not None:                                                 # :364-367 (HR outputs vs the SISR image)
        losses["coarse_mse_sr"] = torch.nn.functional.mse_loss(c_rgb_ori, target_sr)
        losses["fine_mse_sr"] = torch.nn.functional.mse_loss(f_rgb_ori, target_sr)
        tot = tot + (losses["coarse_mse_sr"] + losses["fine_mse_sr"])
    if out_ref is not None:                                                   # :369-372
        losses["ref_coarse_mse"] = torch.nn.functional.mse_loss(out_ref["coarse_comp_rgbs"], ref_rgbs) / (s ** 2)
        losses["ref_fine_mse"] = torch.nn.functional.mse_loss(out_ref["fine_comp_rgbs"], ref_rgbs) / (s ** 2)
        tot = tot + (losses["ref_coarse_mse"] + losses["ref_fine_mse"])
    if tcfg.use_var_loss:                                                     # :374-375
        tot = tot + (tcfg.lambda_coarse_var * losses["coarse_var"] + tcfg.lambda_fine_var * losses["fine_var"])
    if tcfg.use_depth_var_loss:                                               # :376-378
        tot = tot + (tcfg.lambda_coarse_depth_var * losses["coarse_depth_var"]
                     + tcfg.lambda_fine_depth_var * losses["fine_depth_var"])
    losses["tot"] = tot
    tot.backward()                                                            # :396
    with torch.no_grad():                                                     # :379-384
        losses["coarse_psnr"] = -10 * torch.log10(torch.mean((lr_c - target_lr) ** 2))
        if cfg.N_importance > 0:
            losses["fine_psnr"] = -10 * torch.log10(torch.mean((lr_f - target_lr) ** 2))
    gc = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in pc_r.items()}
    gf = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in pf_r.items()}
    return ({k: v.detach() for k, v in losses.items()}, gc, gf, {k: v.detach() for k, v in out.items()})


def subpixel_variance_sum(x: Tensor, n_lr: int, s: int, divisor=None) -> Tensor:
    """``torch.sum(torch.var(torch.reshape(x, (n_lr, s*s, -1)) [/ far], dim=1))`` -- unbiased variance over the
    s*s sub-pixel rays of each LR pixel, summed over pixels and channels (models/nerf_downX_model.py:331-335,
    349-353).  ``divisor`` is the reference's ``self.far``: there a shape-(1,) float32 numpy array (:284), which torch
    refuses to divide a grad-requiring tensor by; the value as a Python float is what the expression means (and what
    oracle/make_golden_train.py hands the reference model when it pins this term)."""
    y = torch.reshape(x, (n_lr, s * s, -1))
    if divisor is not None:
        y = y / divisor
    return torch.sum(torch.var(y, dim=1))


def clip_grads(grads: List[Tensor], tcfg: TrainConfig) -> Optional[Tensor]:
    """nn.utils.clip_grad_norm_ (2-norm over all tensors, coefficient clamp(max/(norm+1e-6), max=1))
    or clip_grad_value_, in place.  Returns the total norm (norm mode)."""
    if tcfg.grad_clip_val <= 0:
        return None
    if tcfg.grad_clip_type == "norm":
        total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g, 2.0) for g in grads]), 2.0)
        coef = torch.clamp(tcfg.grad_clip_val / (total + 1e-6), max=1.0)
        for g in grads:
            g.mul_(coef)
        return total
    for g in grads:
        g.clamp_(min=-tcfg.grad_clip_val, max=tcfg.grad_clip_val)
    return None


def adam_step(params: List[Tensor], grads: List[Tensor], exp_avg: List[Tensor], exp_avg_sq: List[Tensor],
              step: int, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam (single-tensor path, no amsgrad / weight decay / maximize), in place.
    ``step`` is the 1-based step count AFTER the increment."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    bc2_sqrt = bc2 ** 0.5
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)


def lr_at_epoch(tcfg: TrainConfig, epoch: int) -> float:
    """LambdaLR rules of models/networks.py:102-113 (lr after `epoch` scheduler steps)."""
    t = max(0, epoch + 1 - tcfg.n_epochs + tcfg.n_epochs_decay) / float(tcfg.n_epochs_decay + 1)
    if tcfg.lr_policy == "linear":
        lr = tcfg.lr * (1 - t) + tcfg.lr_final * t
    elif tcfg.lr_policy == "exp":
        lr = math.exp(math.log(tcfg.lr) * (1 - t) + math.log(tcfg.lr_final) * t)
    else:
        raise ValueError(tcfg.lr_policy)
    return (lr / tcfg.lr) * tcfg.lr


def golden_sub_indices(numel: int, tag: int, n_sub: int = 256):
    """Fixed pseudo-random subset of a tensor's entries stored in tests/golden/train_step_*.npz."""
    import numpy as np
    g = np.random.Generator(np.random.PCG64(1000003 * tag + numel))
    return np.sort(g.integers(0, numel, size=min(n_sub, numel)))


class TrainState:
    """Parameters + Adam moments of both nets, in the reference's optimiser order
    (itertools.chain(netCoarse.parameters(), netFine.parameters()), :201-203)."""

    def __init__(self, pc: Dict[str, Tensor], pf: Dict[str, Tensor]):
        self.pc = {k: v.detach().clone() for k, v in pc.items()}
        self.pf = {k: v.detach().clone() for k, v in pf.items()}
        self.m = [torch.zeros_like(v) for v in self.param_list()]
        self.v = [torch.zeros_like(v) for v in self.param_list()]
        self.step = 0

    def param_list(self) -> List[Tensor]:
        return list(self.pc.values()) + list(self.pf.values())


def optimize_parameters(state: TrainState, rays: Tensor, target_lr: Tensor, cfg: O.RenderConfig,
                        tcfg: TrainConfig, rng: Optional[O.RenderRng], s: int = 2, lr: Optional[float] = None,
                        z_fine_override: Optional[Tensor] = None, target_sr: Optional[Tensor] = None,
                        ref_rays: Optional[Tensor] = None, ref_rgbs: Optional[Tensor] = None,
                        ref_rng: Optional[O.RenderRng] = None):
    """One full reference training iteration on ``state`` (in place).  Returns (losses, grads list)."""
    losses, gc, gf, _ = loss_and_grads(state.pc, state.pf, rays, target_lr, cfg, tcfg, rng, s, z_fine_override,
                                       target_sr=target_sr, ref_rays=ref_rays, ref_rgbs=ref_rgbs, ref_rng=ref_rng)
    grads = [gc[k] for k in state.pc] + [gf[k] for k in state.pf]
    clip_grads(grads, tcfg)
    state.step += 1
    adam_step(state.param_list(), grads, state.m, state.v, state.step, tcfg.lr if lr is None else lr,
              tcfg.beta1, tcfg.beta2, tcfg.eps)
    return losses, grads
