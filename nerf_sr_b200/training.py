"""Host-side mirror of the reference's training iteration around the render hot path (scope row f-1).

    Renderer-level seams (thin wrappers over the C ABI, include/nsr.h "training" section):
        render_train / backward / lr_loss_grad / clip_coef / adam_step
    RenderFunction            torch.autograd.Function: forward_rays with autograd-connected outputs, so the
                              reference's ``loss_tot.backward(); optimizer.step()`` run unchanged
                              (models/nerf_downX_model.py:390-408)
    Trainer.optimize_parameters   the whole reference iteration on the device: forward (train mode) ->
                              box average + ColorMSELoss (+PSNR, + the sub-pixel variance / SISR terms) -> backward -> [gradient all-reduce] ->
                              clip -> Adam -> re-pack; no host sync, losses stay on the device.

PyTorch provides device memory, streams, RNG draws (in the reference's order) and NCCL only; every
arithmetic step runs in libnsr_b200.  There is no fallback path."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Mapping, Optional, Sequence

import torch

from ._lib import NsrError
from .renderer import Renderer, state_dict_order
from .train_seams import GRAD_KEYS, OUT_KEYS, image_bytes  # noqa: F401  (re-exported: tests and callers import them from here)



def frozen_slices(names: Sequence[str], numels: Sequence[int], fix_layers: Optional[str]) -> List[tuple]:
    """[lo, hi) ranges of the flat gradient that belong to parameters whose name matches the reference's ``--fix_layers``
    regular expression (``re.match`` on ``named_parameters()`` names, models/base_model.py:96-103: those parameters get
    ``requires_grad = False``, so they receive no gradient, stay out of the clipping norm and are skipped by Adam)."""
    import re
    out, off = [], 0
    for n, k in zip(names, numels):
        if fix_layers and re.match(fix_layers, n):
            if out and out[-1][1] == off:
                out[-1] = (out[-1][0], off + int(k))
            else:
                out.append((off, off + int(k)))
        off += int(k)
    return out


def unflatten_grads(flat: torch.Tensor, shapes: Sequence[Sequence[int]]) -> List[torch.Tensor]:
    out, off = [], 0
    for s in shapes:
        n = int(math.prod(s))
        out.append(flat[off:off + n].view(*s))
        off += n
    return out


# ---- autograd bridge ---------------------------------------------------------------------------------
class RenderFunction(torch.autograd.Function):
    """forward_rays whose outputs are connected to the parameters of netCoarse / netFine.

    apply(renderer, rays, rng_dict_or_None, n_coarse_params, ddp_group_or_None, *params) -> the 8 tensors of OUT_KEYS.
    The backward hands dL/d(comp_rgbs, depth, opacity) to the CUDA library and returns per-parameter
    gradients (views of the two flat buffers).  ``ddp_group``: the process group of the DistributedDataParallel
    wrappers around the nets (models/networks.py:72-86).  This function reads the raw parameters, so DDP's own
    ``forward`` -- the only place its reducer is armed -- never runs; the backward therefore does the reducer's job
    itself: one all-reduce (mean) of the two flat gradient buffers over that group."""

    @staticmethod
    def forward(ctx, renderer: Renderer, rays: torch.Tensor, rng, n_coarse: int, ddp_group, *params: torch.Tensor):
        with torch.no_grad():
            renderer.load_params(0, params[:n_coarse])
            renderer.load_params(1, params[n_coarse:])
            ws = renderer.new_train_workspace(rays.shape[0])      # private: several forwards may precede their backwards
            out = renderer.render_train(rays, rng, ws=ws)
        ctx.renderer, ctx.rays, ctx.rng, ctx.ws = renderer, rays, rng, ws
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.n_coarse = n_coarse
        ctx.ddp_group = ddp_group
        ctx.set_materialize_grads(False)
        outs = tuple(out[k] for k in OUT_KEYS)
        ctx.mark_non_differentiable(outs[3], outs[7])
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        grads = dict(zip(OUT_KEYS, gouts))
        gc, gf = ctx.renderer.backward(ctx.rays, ctx.rng, grads, ws=ctx.ws)
        ctx.ws = None                                             # release the stash
        if ctx.ddp_group is not None:
            from .parallel import allreduce_mean_
            allreduce_mean_([gc, gf], ctx.ddp_group)
        nc = ctx.n_coarse
        pg = unflatten_grads(gc, ctx.shapes[:nc]) + unflatten_grads(gf, ctx.shapes[nc:])
        return (None, None, None, None, None, *pg)




def module_params_in_order(net) -> List[torch.Tensor]:
    """Parameters of a reference VanillaMLP (optionally DP/DDP-wrapped) in state_dict order."""
    m = net.module if hasattr(net, "module") else net
    named = dict(m.named_parameters())
    D = sum(1 for k in named if k.startswith("xyz_encoding_") and k.endswith(".0.weight"))
    return [named[n] for n in state_dict_order(D)]


def trainer_kwargs_from_opt(opt) -> Dict[str, object]:
    """The ``Trainer`` arguments for a reference option namespace (options/train_options.py:28-55,
    models/nerf_model.py lambda flags, models/nerf_downX_model.py:107-112): learning rate and Adam beta1, loss weights
    (the variance weights count only when their ``--use_*`` switch is on), clipping, ``--downscale``, ``--fix_layers``."""
    g = lambda name, default: getattr(opt, name, default)
    return dict(
        lr=g("lr", 5e-4), beta1=g("beta1", 0.9),
        lambda_coarse_mse=g("lambda_coarse_mse", 1.0), lambda_fine_mse=g("lambda_fine_mse", 1.0),
        grad_clip_val=g("grad_clip_val", 0.0) or 0.0, grad_clip_type=g("grad_clip_type", "norm"),
        downscale=g("downscale", 1),
        lambda_coarse_var=g("lambda_coarse_var", 0.01) if g("use_var_loss", False) else 0.0,
        lambda_fine_var=g("lambda_fine_var", 0.01) if g("use_var_loss", False) else 0.0,
        lambda_coarse_depth_var=g("lambda_coarse_depth_var", 0.01) if g("use_depth_var_loss", False) else 0.0,
        lambda_fine_depth_var=g("lambda_fine_depth_var", 0.01) if g("use_depth_var_loss", False) else 0.0,
        fix_layers=g("fix_layers", None))


# ---- the fused training iteration -----------------------------------------------------------------------
class Trainer:
    """The reference's ``optimize_parameters`` (models/nerf_downX_model.py:398-408) on the device.

    Owns fp32 master parameters of both nets (state_dict order) and the Adam moments as flat buffers;
    ``optimize_parameters(rays, target_lr)`` runs forward (train mode), the LR loss, the backward, an
    optional gradient all-reduce (DDP semantics: mean over ranks), clipping, Adam and the re-pack of
    the tensor-core weight images.  Nothing synchronises with the host; ``last_metrics`` is a device
    tensor [coarse lam*mse, coarse psnr, fine lam*mse, fine psnr]."""

    def __init__(self, renderer: Renderer, params_coarse: Mapping[str, torch.Tensor], params_fine: Mapping[str, torch.Tensor],
                 lr: float = 5e-4, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, lambda_coarse_mse: float = 1.0,
                 lambda_fine_mse: float = 1.0, grad_clip_val: float = 0.0, grad_clip_type: str = "norm", downscale: int = 2,
                 group=None, lambda_coarse_var: float = 0.0, lambda_fine_var: float = 0.0, lambda_coarse_depth_var: float = 0.0,
                 lambda_fine_depth_var: float = 0.0, fix_layers: Optional[str] = None, allreduce: str = "auto"):
        """allreduce: how the gradient bucket is averaged over the ranks of ``group`` (DDP semantics, models/networks.py:72-86):
        "p2p" = libnsr_b200's own one-kernel all-reduce over peer-mapped memory (nsr_comm_allreduce_mean: NVLink loads /
        stores, fixed rank order -> bit-identical on every rank), "nccl" = one ncclAvg all-reduce on the flat bucket,
        "auto" = p2p on CUDA with NCCL, else the torch.distributed fallback (gloo in the CPU tests).
        lambda_*_var / lambda_*_depth_var: the reference's ``--lambda_*`` values when ``--use_var_loss`` /
        ``--use_depth_var_loss`` are given (models/nerf_downX_model.py:107-112), 0 (default) otherwise.
        fix_layers: the reference's ``--fix_layers`` regex; matching parameters of both nets are frozen: their gradient
        slices are zeroed before the all-reduce / clipping / Adam, which leaves parameter and moments untouched (Adam with
        an identically zero gradient history is the identity) -- the effect of ``requires_grad = False``."""
        self.r = renderer
        dev = renderer.device
        names = state_dict_order(renderer.cfg.D)
        take = lambda sd: [sd[n].detach().to(dev, torch.float32).contiguous().clone() for n in names]
        self.names = names
        self.params = [take(params_coarse), take(params_fine)]
        numel = int(renderer.lib.nsr_grad_numel(renderer._h))
        self.m = [torch.zeros(numel, device=dev), torch.zeros(numel, device=dev)]
        self.v = [torch.zeros(numel, device=dev), torch.zeros(numel, device=dev)]
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.lam = (lambda_coarse_mse, lambda_fine_mse)
        self.lam_var = (lambda_coarse_var, lambda_fine_var)
        self.lam_dvar = (lambda_coarse_depth_var, lambda_fine_depth_var)
        self.last_terms: Optional[torch.Tensor] = None
        self.last_ref_terms: Optional[torch.Tensor] = None
        self.clip_val, self.clip_type = grad_clip_val, grad_clip_type
        self.s = downscale
        self.step = 0
        self.group = group
        self.last_metrics: Optional[torch.Tensor] = None
        self.last_grads = None
        # ONE flat gradient bucket [coarse | fine] (2 x 595 844 fp32 = 4.77 MB): nsr_backward's reduction kernel writes the
        # two halves in place, the data-parallel all-reduce works on the whole bucket (no cat / copy-back launches).  With
        # more than one rank the bucket lives in the reducer's symmetric (peer-mapped) memory.
        self._numel = numel
        self._reducer = None
        self.allreduce_impl = None
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size(group)
        if world > 1:
            from .parallel import make_grad_reducer
            self._reducer = make_grad_reducer(renderer, 2 * numel, group, allreduce)
            self.allreduce_impl = self._reducer.impl
            self._gflat = self._reducer.buffer
        else:
            self._gflat = torch.empty(2 * numel, device=dev, dtype=torch.float32)
        self._frozen = frozen_slices(names, [int(renderer.lib.nsr_param_numel(renderer._h, i)) for i in range(len(names))], fix_layers)
        for w in (0, 1):
            renderer.load_params(w, self.params[w])

    def state_dict(self, which: int) -> Dict[str, torch.Tensor]:
        return dict(zip(self.names, self.params[which]))

    def save_networks(self, save_dir: str, epoch) -> None:
        """``<epoch>_net_Coarse.pth`` / ``<epoch>_net_Fine.pth`` in the reference's format (models/base_model.py:181-196)."""
        from .checkpoints import save_networks
        save_networks(save_dir, epoch, self.state_dict(0), self.state_dict(1))

    def load_networks(self, save_dir: str, epoch, keys: Optional[str] = None) -> None:
        """Overwrite the master parameters from a checkpoint written by this class or by the reference
        (models/base_model.py:198-219; ``keys``: ``--init_weights_keys``) and re-pack the tensor-core weight images.
        Adam moments and the step count are kept (the reference does not checkpoint them either)."""
        from .checkpoints import load_networks
        for w, sd in enumerate(load_networks(save_dir, epoch, keys)):
            have = dict(zip(self.names, self.params[w]))
            for k, v in sd.items():
                if k not in have:
                    raise NsrError(1, f"unexpected key {k!r} in checkpoint")
                if tuple(v.shape) != tuple(have[k].shape):
                    raise NsrError(1, f"{k}: checkpoint shape {tuple(v.shape)} != {tuple(have[k].shape)}")
                have[k].copy_(v.to(have[k].device, torch.float32))
            if keys is None and len(sd) != len(have):
                raise NsrError(1, f"checkpoint has {len(sd)} tensors, the architecture {len(have)}")
            self.r.load_params(w, self.params[w])

    def draw_rng(self, n_rays: int, generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
        """The reference's train-mode draws, in its order (models/utils.py:41, :210, :73, :210)."""
        c = self.r.cfg
        dev = self.r.device
        rng = {"u_coarse": torch.rand(n_rays, c.n_coarse, device=dev, generator=generator)}
        if c.noise_std > 0:
            rng["noise_coarse"] = torch.randn(n_rays, c.n_coarse, device=dev, generator=generator)
        rng["u_fine"] = torch.rand(n_rays, c.n_importance, device=dev, generator=generator)
        if c.noise_std > 0:
            rng["noise_fine"] = torch.randn(n_rays, c.n_coarse + c.n_importance, device=dev, generator=generator)
        return rng

    def grad_views(self):
        """(grad_coarse, grad_fine): the two halves of the flat all-reduce bucket."""
        return self._gflat[: self._numel], self._gflat[self._numel: 2 * self._numel]

    def allreduce_grads(self, gc: torch.Tensor, gf: torch.Tensor) -> None:
        """Average (gc, gf) over the ranks in place.  One collective on the flat bucket when they are its halves."""
        if self._reducer is None:
            return
        a, b = self.grad_views()
        if gc.data_ptr() == a.data_ptr() and gf.data_ptr() == b.data_ptr():
            self._reducer.allreduce_mean_()
        else:
            from .parallel import allreduce_mean_
            allreduce_mean_([gc, gf], self.group)

    def forward_backward(self, rays: torch.Tensor, target_lr: torch.Tensor, rng=None, target_sr: Optional[torch.Tensor] = None,
                         far: Optional[float] = None, ref_rays: Optional[torch.Tensor] = None,
                         ref_rgbs: Optional[torch.Tensor] = None, ref_rng=None, out: Optional[Sequence[torch.Tensor]] = None):
        """forward + loss + backward; returns (grad_coarse_flat, grad_fine_flat), sets last_metrics.
        ``target_sr`` [N,3]: the SISR supervision ``data_rgbs_sr`` (``--sisr_path``).  ``far``: the reference's
        ``self.far`` for the depth-variance term (default: read from ``rays[0, 7]`` like the reference -- one host sync;
        datasets have a constant far plane, pass it to stay asynchronous).  With any of the extra terms on,
        ``last_terms`` [2,8] holds (lambda*mse, psnr, var_sum, depth_var_sum, mse_sr, total, 0, 0) per net.
        ``ref_rays`` [M,8] / ``ref_rgbs`` [M,3] / ``ref_rng``: the reference-view batch of ``--with_ref``
        (models/nerf_downX_model.py:321-324, 369-372): a second forward (its train-mode draws come after the main
        batch's) whose HR colours enter the loss as MSE / s^2; ``last_ref_terms`` [2] holds the two terms."""
        r = self.r
        if ref_rays is not None:
            return self._forward_backward_with_ref(rays, target_lr, rng, target_sr, far, ref_rays, ref_rgbs, ref_rng, out)
        o = r.render_train(rays, rng, want_weights=False)
        if target_sr is not None or any(self.lam_var) or any(self.lam_dvar):
            return r.backward(rays, rng, self._main_loss_grads(o, rays, target_lr, target_sr, far), out=out)
        m = torch.empty(4, device=r.device, dtype=torch.float32)
        _, _, g_c = r.lr_loss_grad(o["coarse_comp_rgbs"], target_lr, self.s, self.lam[0], metrics_out=m[0:2])
        _, _, g_f = r.lr_loss_grad(o["fine_comp_rgbs"], target_lr, self.s, self.lam[1], metrics_out=m[2:4])
        self.last_metrics = m
        return r.backward(rays, rng, {"coarse_comp_rgbs": g_c, "fine_comp_rgbs": g_f}, out=out)

    def _main_loss_grads(self, out, rays, target_lr, target_sr, far):
        """Loss terms + dL/d(outputs) of the main batch through nsr_loss_epilogue."""
        r = self.r
        if any(self.lam_dvar) and far is None:
            far = float(rays[0, 7])
        terms, grads = [], {}
        for w, net in enumerate(("coarse", "fine")):
            e = r.loss_epilogue(out[f"{net}_comp_rgbs"], target_lr, self.s, self.lam[w],
                                hr_depth=out[f"{net}_depth"] if self.lam_dvar[w] else None, lambda_var=self.lam_var[w],
                                lambda_depth_var=self.lam_dvar[w], far=far or 0.0, target_hr=target_sr)
            terms.append(e["metrics"])
            grads[f"{net}_comp_rgbs"] = e["g_rgb"]
            if "g_depth" in e:
                grads[f"{net}_depth"] = e["g_depth"]
        self.last_terms = torch.stack(terms)
        self.last_metrics = torch.cat([terms[0][:2], terms[1][:2]])
        return grads

    def _forward_backward_with_ref(self, rays, target_lr, rng, target_sr, far, ref_rays, ref_rgbs, ref_rng, grad_out=None):
        r = self.r
        if ref_rgbs is None or ref_rays.shape[0] != ref_rgbs.shape[0] or ref_rays.shape[0] % (self.s * self.s):
            raise NsrError(1, "ref_rays / ref_rgbs must have the same number of rows, a multiple of downscale^2")
        ws_main = r.new_train_workspace(rays.shape[0])           # two forwards are in flight before the first backward
        ws_ref = r.new_train_workspace(ref_rays.shape[0])
        out = r.render_train(rays, rng, want_weights=False, ws=ws_main)
        out_ref = r.render_train(ref_rays, ref_rng, want_weights=False, ws=ws_ref)
        grads = self._main_loss_grads(out, rays, target_lr, target_sr, far)
        ref_terms, ref_grads = [], {}
        for net in ("coarse", "fine"):                            # mse(out_ref_*_comp_rgbs, data_ref_rgbs) / s^2
            e = r.loss_epilogue(out_ref[f"{net}_comp_rgbs"], None, self.s, 0.0, target_hr=ref_rgbs, lambda_hr=1.0 / (self.s * self.s))
            ref_terms.append(e["metrics"][4])
            ref_grads[f"{net}_comp_rgbs"] = e["g_rgb"]
        self.last_ref_terms = torch.stack(ref_terms)
        gc, gf = r.backward(rays, rng, grads, ws=ws_main, out=grad_out)
        gc2, gf2 = r.backward(ref_rays, ref_rng, ref_grads, ws=ws_ref)
        return gc.add_(gc2), gf.add_(gf2)                        # autograd's accumulation over the two forward graphs

    def optimize_parameters(self, rays: torch.Tensor, target_lr: torch.Tensor, rng=None, lr: Optional[float] = None,
                            target_sr: Optional[torch.Tensor] = None, far: Optional[float] = None,
                            ref_rays: Optional[torch.Tensor] = None, ref_rgbs: Optional[torch.Tensor] = None, ref_rng=None):
        r = self.r
        gc, gf = self.forward_backward(rays, target_lr, rng, target_sr=target_sr, far=far, ref_rays=ref_rays, ref_rgbs=ref_rgbs,
                                       ref_rng=ref_rng, out=self.grad_views())
        for lo, hi in self._frozen:                    # --fix_layers
            gc[lo:hi].zero_()
            gf[lo:hi].zero_()
        self.allreduce_grads(gc, gf)                    # one 4.77 MB bucket, one kernel (DDP: models/networks.py:72-86)
        coef, clip_value = None, 0.0
        if self.clip_val > 0:
            if self.clip_type == "norm":
                coef = r.clip_coef(gc, gf, self.clip_val)
            else:
                clip_value = self.clip_val
        self.step += 1
        for w, g in ((0, gc), (1, gf)):
            r.adam_step(self.params[w], g, self.m[w], self.v[w], self.step, self.lr if lr is None else lr, self.beta1,
                        self.beta2, self.eps, coef, clip_value)
            r.load_params(w, self.params[w])
        self.last_grads = (gc, gf)
        return self.last_metrics
