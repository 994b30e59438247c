"""The training seams of the C ABI (include/nsr.h, "training" section) as methods of ``Renderer``: thin wrappers that
allocate outputs, marshal pointers and check status.  ``Renderer`` (renderer.py) inherits this class; nothing here
imports it, so there is no import cycle and no method is attached after the fact.

    render_train / backward / lr_loss_grad / loss_epilogue / clip_coef / adam_step / load_params   the iteration's seams
    pack_image / unpack_image / relu_bits / stash_* / debug_dx / debug_dw / train_layout            test seams

PyTorch provides device memory and streams only; every arithmetic step runs in libnsr_b200.  There is no fallback path."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional, Sequence

import torch

from ._lib import NsrError, NsrLossTerms, NsrOutGrads, NsrOutputs, NsrRng

OUT_KEYS = ("coarse_comp_rgbs", "coarse_depth", "coarse_opacity", "coarse_weights",
            "fine_comp_rgbs", "fine_depth", "fine_opacity", "fine_weights")
GRAD_KEYS = ("coarse_comp_rgbs", "coarse_depth", "coarse_opacity", "fine_comp_rgbs", "fine_depth", "fine_opacity")


def _rng_struct(r, rng: Optional[Mapping[str, torch.Tensor]], keep: list) -> Optional[NsrRng]:
    if rng is None:
        return None
    s = NsrRng()
    for k in ("u_coarse", "noise_coarse", "u_fine", "noise_fine"):
        t = rng.get(k) if isinstance(rng, Mapping) else getattr(rng, k, None)
        if t is not None:
            t = r._f32(t.to(r.device))
            keep.append(t)
            setattr(s, k, t.data_ptr())
    return s



def image_bytes(n_rows: int, n_cols: int) -> int:
    return ((n_rows + 127) // 128) * (n_cols // 64) * 32768



class TrainSeams:
    """Mixed into ``Renderer``; expects ``self.lib``, ``self._h``, ``self.device``, ``self._check``, ``self._f32``, ``self._stream``,
    ``self._keep`` and the sample counts of the handle."""

    def _train_workspace(self, n_rays: int) -> torch.Tensor:
        need = self.lib.nsr_train_workspace_bytes(self._h, n_rays)
        ws = getattr(self, "_train_ws", None)
        if ws is None or ws.numel() < need:
            self._train_ws = None
            ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._train_ws = ws
        return ws


    def new_train_workspace(self, n_rays: int) -> torch.Tensor:
        """A private stash buffer for one forward/backward pair (several may be in flight at once)."""
        return torch.empty(self.lib.nsr_train_workspace_bytes(self._h, n_rays), dtype=torch.uint8, device=self.device)


    def render_train(self, rays: torch.Tensor, rng=None, want_weights: bool = True, want_z_fine: bool = False,
                     ws: Optional[torch.Tensor] = None):
        """forward_rays in train mode, keeping the activation stash for ``backward`` (same outputs as
        forward_rays).  The stash goes to ``ws`` (new_train_workspace) or, by default, to a renderer-owned buffer that the
        next render_train overwrites: pass the same ``ws`` to ``backward``."""
        rays = self._f32(rays)
        n, stride = rays.shape
        dev, f32 = self.device, torch.float32
        out = {"coarse_comp_rgbs": torch.empty(n, 3, device=dev, dtype=f32), "coarse_depth": torch.empty(n, device=dev, dtype=f32),
               "coarse_opacity": torch.empty(n, device=dev, dtype=f32),
               "fine_comp_rgbs": torch.empty(n, 3, device=dev, dtype=f32), "fine_depth": torch.empty(n, device=dev, dtype=f32),
               "fine_opacity": torch.empty(n, device=dev, dtype=f32)}
        if want_weights:
            out["coarse_weights"] = torch.empty(n, self.n_coarse, device=dev, dtype=f32)
            out["fine_weights"] = torch.empty(n, self.n_fine, device=dev, dtype=f32)
        if want_z_fine:
            out["z_fine"] = torch.empty(n, self.n_fine, device=dev, dtype=f32)
        o = NsrOutputs()
        for k, v in out.items():
            setattr(o, k, v.data_ptr())
        keep: list = []
        r = _rng_struct(self, rng, keep)
        if ws is None:
            ws = self._train_workspace(n)
        self._check(self.lib.nsr_render_train(self._h, rays.data_ptr(), n, stride, C.byref(r) if r is not None else None,
                                              C.byref(o), ws.data_ptr(), ws.numel(), self._stream()))
        return out


    def backward(self, rays: torch.Tensor, rng, grads: Mapping[str, Optional[torch.Tensor]],
                 ws: Optional[torch.Tensor] = None, out: Optional[Sequence[torch.Tensor]] = None):
        """dL/d(outputs) -> (grad_coarse_flat, grad_fine_flat): flat fp32 gradients in state_dict order.
        ``ws``: the stash buffer the matching render_train filled (default: the renderer-owned one).
        ``out``: (grad_coarse, grad_fine) buffers to fill (e.g. the two halves of one flat all-reduce bucket)."""
        rays = self._f32(rays)
        n, stride = rays.shape
        g = NsrOutGrads()
        keep: list = []
        for k in GRAD_KEYS:
            t = grads.get(k)
            if t is not None:
                t = self._f32(t).reshape(n, -1)
                keep.append(t)
                setattr(g, k, t.data_ptr())
        for k in ("coarse_weights", "fine_weights"):
            if grads.get(k) is not None:
                raise NsrError(2, f"gradient w.r.t. {k} is not supported (the reference's losses never use it)")
        r = _rng_struct(self, rng, keep)
        numel = int(self.lib.nsr_grad_numel(self._h))
        if out is not None:
            gc, gf = out
            for t in (gc, gf):
                if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != numel or t.device != self.device:
                    raise NsrError(1, f"backward(out=...) needs two contiguous fp32 buffers of {numel} elements on {self.device}")
        else:
            gc = torch.empty(numel, device=self.device, dtype=torch.float32)
            gf = torch.empty(numel, device=self.device, dtype=torch.float32)
        if ws is None:
            ws = self._train_workspace(n)
        self._check(self.lib.nsr_backward(self._h, rays.data_ptr(), n, stride, C.byref(r) if r is not None else None, C.byref(g),
                                          gc.data_ptr(), gf.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()))
        return gc, gf


    def lr_loss_grad(self, hr_rgb: torch.Tensor, target_lr: torch.Tensor, s: int, lam: float = 1.0,
                     want_grad: bool = True, metrics_out: Optional[torch.Tensor] = None):
        """(lr_rgb [n_lr,3], metrics [2] = (lam*mse, psnr), g_hr [n_lr*s*s,3] = d(lam*mse)/d(hr_rgb))."""
        hr_rgb, target_lr = self._f32(hr_rgb), self._f32(target_lr)
        n_lr = target_lr.shape[0]
        if hr_rgb.shape[0] != n_lr * s * s:
            raise NsrError(1, f"hr_rgb has {hr_rgb.shape[0]} rows, expected {n_lr}*{s}*{s}")
        lr = torch.empty(n_lr, 3, device=self.device, dtype=torch.float32)
        m = metrics_out if metrics_out is not None else torch.empty(2, device=self.device, dtype=torch.float32)
        g = torch.empty_like(hr_rgb) if want_grad else None
        self._check(self.lib.nsr_lr_loss_grad(self._h, hr_rgb.data_ptr(), target_lr.data_ptr(), n_lr, s, float(lam), lr.data_ptr(),
                                              m.data_ptr(), g.data_ptr() if g is not None else None, self._stream()))
        return lr, m, g


    def loss_epilogue(self, hr_rgb: torch.Tensor, target_lr: Optional[torch.Tensor], s: int, lambda_mse: float = 1.0,
                      hr_depth: Optional[torch.Tensor] = None, lambda_var: float = 0.0, lambda_depth_var: float = 0.0,
                      far: float = 0.0, target_hr: Optional[torch.Tensor] = None, want_grad: bool = True,
                      lambda_hr: float = 1.0) -> Dict[str, torch.Tensor]:
        """Every term of the reference's ``calculate_losses`` for one net's outputs (models/nerf_downX_model.py:326-378)
        and its gradient down to the HR outputs, in one launch (nsr_loss_epilogue).  ``lambda_var`` /
        ``lambda_depth_var`` = 0 switch the sub-pixel variance terms off (``--use_var_loss`` / ``--use_depth_var_loss``
        not given); ``target_hr`` is ``data_rgbs_sr`` (``--sisr_path``, ``lambda_hr`` 1) or, with ``target_lr`` None, the
        reference-view colours ``data_ref_rgbs`` (``--with_ref``, ``lambda_hr`` 1/s^2); ``far`` is the reference's ``self.far``.
        Returns lr_rgb [n_lr,3], lr_depth [n_lr] (if hr_depth), metrics [8] = (lambda*mse, psnr, var_sum,
        depth_var_sum, lambda_hr*mse_hr, total, 0, 0), g_rgb, g_depth (if want_grad)."""
        hr_rgb = self._f32(hr_rgb)
        if hr_rgb.shape[0] % (s * s):
            raise NsrError(1, f"hr_rgb has {hr_rgb.shape[0]} rows, not a multiple of {s}*{s}")
        n_lr = hr_rgb.shape[0] // (s * s)
        n = n_lr * s * s
        if target_lr is not None:
            target_lr = self._f32(target_lr)
            if tuple(target_lr.shape) != (n_lr, 3):
                raise NsrError(1, f"target_lr is {tuple(target_lr.shape)}, expected ({n_lr}, 3) for {hr_rgb.shape[0]} HR rows at s={s}")
        dev, f32 = self.device, torch.float32
        out = {"lr_rgb": torch.empty(n_lr, 3, device=dev, dtype=f32), "metrics": torch.empty(8, device=dev, dtype=f32)}
        if hr_depth is not None:
            hr_depth = self._f32(hr_depth).reshape(-1)
            if hr_depth.shape[0] != n:
                raise NsrError(1, f"hr_depth has {hr_depth.shape[0]} rows, expected {n}")
            out["lr_depth"] = torch.empty(n_lr, device=dev, dtype=f32)
        if target_hr is not None:
            target_hr = self._f32(target_hr)
            if tuple(target_hr.shape) != (n, 3):
                raise NsrError(1, f"target_hr is {tuple(target_hr.shape)}, expected ({n}, 3)")
        if want_grad:
            out["g_rgb"] = torch.empty(n, 3, device=dev, dtype=f32)
            if hr_depth is not None:
                out["g_depth"] = torch.empty(n, device=dev, dtype=f32)
        t = NsrLossTerms()
        t.struct_size = C.sizeof(NsrLossTerms)
        t.s, t.lambda_mse, t.lambda_var, t.lambda_depth_var, t.far_plane = int(s), float(lambda_mse), float(lambda_var), \
            float(lambda_depth_var), float(far)
        t.lambda_hr = float(lambda_hr)
        ptr = lambda k: out[k].data_ptr() if k in out else None
        self._check(self.lib.nsr_loss_epilogue(self._h, hr_rgb.data_ptr(), hr_depth.data_ptr() if hr_depth is not None else None,
                                               target_lr.data_ptr() if target_lr is not None else None,
                                               target_hr.data_ptr() if target_hr is not None else None,
                                               n_lr, C.byref(t), ptr("lr_rgb"), ptr("lr_depth"), ptr("metrics"), ptr("g_rgb"),
                                               ptr("g_depth"), self._stream()))
        return out


    def clip_coef(self, grad_a: torch.Tensor, grad_b: Optional[torch.Tensor], max_norm: float) -> torch.Tensor:
        out = torch.empty(2, device=self.device, dtype=torch.float32)
        self._check(self.lib.nsr_clip_coef(self._h, grad_a.data_ptr(), grad_b.data_ptr() if grad_b is not None else None,
                                           grad_a.numel(), float(max_norm), out.data_ptr(), self._stream()))
        return out


    def adam_step(self, params: Sequence[torch.Tensor], grad_flat: torch.Tensor, exp_avg: torch.Tensor,
                  exp_avg_sq: torch.Tensor, step: int, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                  clip_coef_dev: Optional[torch.Tensor] = None, clip_value: float = 0.0):
        arr, _ = self._checked_param_array(params, "adam_step")
        self._check(self.lib.nsr_adam_step(self._h, arr, len(params), grad_flat.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                           int(step), float(lr), float(beta1), float(beta2), float(eps),
                                           clip_coef_dev.data_ptr() if clip_coef_dev is not None else None, float(clip_value),
                                           self._stream()))


    # ---- test seams of the backward GEMMs ----
    def pack_image(self, x: torch.Tensor, n_cols: Optional[int] = None) -> torch.Tensor:
        x = self._f32(x)
        n_rows, ld = x.shape
        n_cols = n_cols or ((ld + 63) // 64) * 64
        img = torch.empty(image_bytes(n_rows, n_cols), dtype=torch.uint8, device=self.device)
        self._check(self.lib.nsr_debug_pack_image(self._h, x.data_ptr(), n_rows, n_cols, ld, img.data_ptr(), self._stream()))
        return img


    def unpack_image(self, img: torch.Tensor, n_rows: int, n_cols: int, ld: Optional[int] = None) -> torch.Tensor:
        ld = ld or n_cols
        out = torch.empty(n_rows, ld, device=self.device, dtype=torch.float32)
        self._check(self.lib.nsr_debug_unpack_image(self._h, img.data_ptr(), n_rows, n_cols, ld, out.data_ptr(), self._stream()))
        return out


    def relu_bits(self, x: torch.Tensor) -> torch.Tensor:
        """[rows,256] fp32 -> the 1-bit-per-activation ReLU mask layout the forward stashes (test seam)."""
        x = self._f32(x)
        rows = x.shape[0]
        out = torch.empty(((rows + 127) // 128) * 128 * 8, dtype=torch.int32, device=self.device)
        self._check(self.lib.nsr_debug_relu_bits(self._h, x.data_ptr(), rows, out.data_ptr(), self._stream()))
        return out


    def stash_mask(self, n_rays: int, which: int, layer: int) -> torch.Tensor:
        """ReLU mask of h_layer (1..8) of pass `which` after render_train, as a [P,256] bool tensor (test seam)."""
        L = self.train_layout(n_rays)
        ws = self._train_ws
        base = (-ws.data_ptr()) % 256
        S = self.n_fine if which else self.n_coarse
        tiles = L[f"tiles{which}"]
        off = base + L[f"mask{which}"] + (layer - 1) * tiles * 128 * 32
        words = ws[off: off + tiles * 128 * 32].view(torch.int32).view(tiles * 128, 8)
        bits = (words.unsqueeze(-1) >> torch.arange(32, device=ws.device, dtype=torch.int32)) & 1
        return bits.reshape(tiles * 128, 256)[: n_rays * S].bool()


    def debug_dx(self, which: int, layer_idx: int, a_img: torch.Tensor, n_rows: int, mask_bits=None, dsig=None, wsig=None):
        out = torch.empty(image_bytes(n_rows, 256), dtype=torch.uint8, device=self.device)
        self._check(self.lib.nsr_debug_dx(self._h, which, layer_idx, a_img.data_ptr(), out.data_ptr(),
                                          mask_bits.data_ptr() if mask_bits is not None else None,
                                          dsig.data_ptr() if dsig is not None else None,
                                          wsig.data_ptr() if wsig is not None else None, n_rows, self._stream()))
        return out


    def debug_dw(self, a_img: torch.Tensor, a_cols: int, blk0: int, blk1: int, b_img: torch.Tensor, b_cols: int, n_rows: int):
        out = torch.empty(128, b_cols, device=self.device, dtype=torch.float32)
        bias = torch.empty(128, device=self.device, dtype=torch.float32)
        scratch = torch.empty(148 * 128 * (b_cols + 1) * 4 + 1024, dtype=torch.uint8, device=self.device)
        self._check(self.lib.nsr_debug_dw(self._h, a_img.data_ptr(), a_cols, blk0, blk1, b_img.data_ptr(), b_cols, out.data_ptr(),
                                          bias.data_ptr(), n_rows, scratch.data_ptr(), scratch.numel(), self._stream()))
        return out, bias


    def train_layout(self, n_rays: int) -> Dict[str, int]:
        """Offsets of the stash regions inside the train workspace (test seam)."""
        arr = (C.c_int64 * 18)()
        self._check(self.lib.nsr_debug_train_layout(self._h, n_rays, arr))
        keys = ["enc0", "h0", "dir0", "raw0", "z0", "tiles0", "enc1", "h1", "dir1", "raw1", "z1", "tiles1",
                "dhead", "dzdir", "g0", "g1", "mask0", "mask1"]
        return dict(zip(keys, [int(x) for x in arr]))


    def stash_activation(self, n_rays: int, which: int, layer: int) -> torch.Tensor:
        """Unpack one stashed tensor of pass `which` after render_train: layer 0 = encoded xyz [P,64],
        1..8 = h_l [P,256], 9 = feat [P,256], 10 = dir activations [P,128] (test seam)."""
        L = self.train_layout(n_rays)
        ws = self._train_ws
        base = (-ws.data_ptr()) % 256
        S = self.n_fine if which else self.n_coarse
        tiles = L[f"tiles{which}"]
        rows = n_rays * S
        if layer == 0:
            off, cols = L[f"enc{which}"], 64
        elif layer <= 9:
            off, cols = L[f"h{which}"] + (layer - 1) * tiles * 4 * 32768, 256
        else:
            off, cols = L[f"dir{which}"], 128
        img = ws[base + off: base + off + tiles * (cols // 64) * 32768]
        return self.unpack_image(img, rows, cols)



    def _checked_param_array(self, params: Sequence[torch.Tensor], what: str):
        """(ctypes pointer array, detached tensors) for a parameter list in state_dict order.  The dtype / layout / size checks
        run once per distinct list of storages (the training loop passes the same tensors every step: ~50 ctypes calls and a
        Python loop per call otherwise, which is first-order once a step is ~1 ms of GPU time)."""
        key = tuple(p.data_ptr() for p in params)
        cache = self.__dict__.setdefault("_param_arrays", {})
        hit = cache.get(key)
        if hit is not None:
            return hit
        ts = [p.detach() for p in params]
        if len(ts) != int(self.lib.nsr_param_count(self._h)):
            raise NsrError(1, f"{what}: expected {int(self.lib.nsr_param_count(self._h))} parameter tensors, got {len(ts)}")
        for i, t in enumerate(ts):
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                raise NsrError(1, f"{what}: parameters must be contiguous fp32 tensors on the renderer's device")
            if t.numel() != self.lib.nsr_param_numel(self._h, i):
                raise NsrError(1, f"{what}: parameter {i}: {tuple(t.shape)} does not match the configured architecture")
        arr = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        if len(cache) > 16:
            cache.clear()
        cache[key] = (arr, ts)                    # `ts` keeps the storages alive, so a data_ptr key cannot be recycled
        return cache[key]


    def load_params(self, which: int, params: Sequence[torch.Tensor]):
        """nsr_pack_weights straight from a parameter list in state_dict order (no name lookup)."""
        arr, ts = self._checked_param_array(params, "load_params")
        self._check(self.lib.nsr_pack_weights(self._h, which, arr, len(ts), self._stream()))
        self._keep[which] = ts


