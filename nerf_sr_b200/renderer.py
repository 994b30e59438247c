"""Host-side mirror of the reference's render interface for the hot path.

``Renderer`` owns one libnsr_b200 handle and exposes the reference's seams with
the same names, argument meaning and result layout:

    forward_rays(rays)            NeRFDownXModel.forward_rays   models/nerf_downX_model.py:280-313
    render_pass(which, rays, z)   render_rays + renderer        :260-278 / models/rendering.py:75-111
    sample_along_rays / resample_along_rays (z-values only)     models/utils.py:17-95
    posenc(x, deg)                PositionalEncoding.__call__   models/embedding.py:44-63
    box_average(x, s)             comp_low_res_output           models/nerf_downX_model.py:337-348
    generate_rays(c2w, ...)       get_ray_directions/get_rays/get_ndc_rays + SS grouping
    render_frame_host(...)        set_input + forward + comp_low_res_output with host buffers

PyTorch is used only for device memory, streams and (in parallel.py) NCCL.  All
arithmetic happens inside the CUDA library; there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional, Sequence

import torch

from . import _lib
from ._lib import NsrConfig, NsrError, NsrLrOutputs, NsrOutputs, NsrPassOutputs, NsrRayGen, NsrRng, PRECISIONS
from .train_seams import TrainSeams


def state_dict_order(D: int = 8):
    """Parameter names in the order nsr_pack_weights expects (models/networks.py:149-180)."""
    names = []
    for i in range(D):
        names += [f"xyz_encoding_{i+1}.0.weight", f"xyz_encoding_{i+1}.0.bias"]
    names += ["xyz_encoding_final.weight", "xyz_encoding_final.bias",
              "dir_encoding.0.weight", "dir_encoding.0.bias",
              "sigma.weight", "sigma.bias", "rgb.0.weight", "rgb.0.bias"]
    return names


def config_from_opt(opt, device_index: int, precision: str = "bf16x3", viewdir_offset: int = 3) -> NsrConfig:
    """Translate the reference's argparse ``opt`` (or the oracle's RenderConfig) into NsrConfig."""
    g = lambda k, d=None: getattr(opt, k, d)
    cfg = NsrConfig()
    cfg.struct_size = C.sizeof(NsrConfig)
    cfg.device = device_index
    cfg.precision = PRECISIONS[precision]
    cfg.D, cfg.W = int(g("D", 8)), int(g("W", 256))
    mask = 0
    for s in (g("skips", (4,)) or ()):
        mask |= 1 << int(s)
    cfg.skips_mask = mask
    cfg.no_dir = int(bool(g("no_dir", False)))
    cfg.color_activation = {"sigmoid": 0, "none": 1}[g("color_activation", "sigmoid")]
    cfg.deg_pos, cfg.deg_dir = int(g("deg_pos", 10)), int(g("deg_dir", 4))
    cfg.no_xyz, cfg.no_logscale = int(bool(g("no_xyz", False))), int(bool(g("no_logscale", False)))
    cfg.n_coarse, cfg.n_importance = int(g("N_coarse", 64)), int(g("N_importance", 64))
    cfg.lindisp, cfg.white_bkgd = int(bool(g("lindisp", False))), int(bool(g("white_bkgd", False)))
    cfg.sigma_activation = {"relu": 0, "softplus": 1}[g("sigma_activation", "relu")]
    cfg.gamma_correct = int(bool(g("gamma_correct", False)))
    cfg.noise_std = float(g("noise_std", 0.0))
    cfg.viewdir_offset = int(g("viewdir_offset", viewdir_offset))
    if int(g("dim_rgb", 3)) != 3 or int(g("dim_pos", 3)) != 3 or int(g("dim_dir", 3)) != 3:
        raise NsrError(2, "dim_rgb/dim_pos/dim_dir other than 3 are not supported")
    return cfg


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Renderer(TrainSeams):
    """One handle = one (opt, device).  Thread-compatible: use one Renderer per thread/stream.
    The training seams (render_train, backward, adam_step, ...) come from ``TrainSeams`` (train_seams.py)."""

    def __init__(self, opt, device: Optional[torch.device] = None, precision: str = "bf16x3",
                 viewdir_offset: int = 3):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise NsrError(6, "no CUDA device: nerf_sr_b200 runs only on the GPU (no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.precision = precision
        self.cfg = config_from_opt(opt, self.device.index or 0, precision, viewdir_offset)
        h = C.c_void_p()
        rc = self.lib.nsr_create(C.byref(self.cfg), C.byref(h))
        if rc != _lib.NSR_OK:
            raise NsrError(rc, self.lib.nsr_last_error(None).decode())
        self._h = h
        self.n_coarse, self.n_importance = self.cfg.n_coarse, self.cfg.n_importance
        self.n_fine = self.n_coarse + self.n_importance
        self._ws: Optional[torch.Tensor] = None
        self._param_versions = [None, None]
        self._keep = [None, None]

    # -- plumbing ----------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.nsr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != _lib.NSR_OK:
            raise NsrError(rc, self.lib.nsr_last_error(self._h).decode())

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _workspace(self, n_rays: int):
        need = self.lib.nsr_workspace_bytes(self._h, n_rays)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws.data_ptr(), self._ws.numel()

    def _f32(self, t: torch.Tensor, cols: Optional[int] = None) -> torch.Tensor:
        if t.device != self.device:
            raise NsrError(1, f"tensor on {t.device}, renderer on {self.device}")
        t = t.detach()
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.float().contiguous()
        return t

    @property
    def launch_count(self) -> int:
        return int(self.lib.nsr_launch_count(self._h))

    def kernel_clock_mhz(self) -> Optional[float]:
        """SM clock (MHz) the most recent fused-pass kernel ran at, from clock64 / globaltimer stamps taken by CTA 0 at
        kernel entry and exit (nsr_debug_kernel_clock).  Synchronises the current stream."""
        v = (C.c_int64 * 4)()
        self._check(self.lib.nsr_debug_kernel_clock(self._h, v, self._stream()))
        dns = v[3] - v[1]
        return None if dns <= 0 else 1e3 * (v[2] - v[0]) / dns

    # -- weights -----------------------------------------------------------------
    def load_state_dict(self, which: int, state_dict: Mapping[str, torch.Tensor]):
        """which: 0 = netCoarse, 1 = netFine.  Accepts the reference's state_dict (optionally with a
        DataParallel/DDP 'module.' prefix, models/base_model.py:193-194)."""
        names = state_dict_order(self.cfg.D)
        tensors = []
        for i, n in enumerate(names):
            t = state_dict[n] if n in state_dict else state_dict["module." + n]
            t = t.detach().to(self.device, torch.float32).contiguous()
            if t.numel() != self.lib.nsr_param_numel(self._h, i):
                raise NsrError(1, f"{n}: {tuple(t.shape)} does not match the configured architecture")
            tensors.append(t)
        arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        self._check(self.lib.nsr_pack_weights(self._h, which, arr, len(tensors), self._stream()))
        self._keep[which] = tensors     # keep alive until the async pack has run

    def sync_from_modules(self, net_coarse, net_fine):
        """Re-pack when parameter versions changed (after every optimiser step / load_networks)."""
        for which, net in enumerate((net_coarse, net_fine)):
            m = net.module if hasattr(net, "module") else net
            ver = tuple(p._version for p in m.parameters()) + tuple(p.data_ptr() for p in m.parameters())
            if ver != self._param_versions[which]:
                self.load_state_dict(which, m.state_dict())
                self._param_versions[which] = ver

    # -- the hot path ------------------------------------------------------------
    def forward_rays(self, rays: torch.Tensor, rng: Optional[Mapping[str, torch.Tensor]] = None,
                     want_weights: bool = True, want_z_fine: bool = False) -> Dict[str, torch.Tensor]:
        """rays [N, 8|11] fp32 on self.device -> the reference's dict of 8 tensors
        (models/nerf_downX_model.py:293-311).  rng: optional dict with u_coarse / noise_coarse /
        u_fine / noise_fine (train mode); None = eval mode."""
        rays = self._f32(rays)
        n, stride = rays.shape
        dev, f32 = self.device, torch.float32
        out = {"coarse_comp_rgbs": torch.empty(n, 3, device=dev, dtype=f32),
               "coarse_depth": torch.empty(n, device=dev, dtype=f32),
               "coarse_opacity": torch.empty(n, device=dev, dtype=f32)}
        if want_weights:
            out["coarse_weights"] = torch.empty(n, self.n_coarse, device=dev, dtype=f32)
        if self.n_importance > 0:
            out["fine_comp_rgbs"] = torch.empty(n, 3, device=dev, dtype=f32)
            out["fine_depth"] = torch.empty(n, device=dev, dtype=f32)
            out["fine_opacity"] = torch.empty(n, device=dev, dtype=f32)
            if want_weights:
                out["fine_weights"] = torch.empty(n, self.n_fine, device=dev, dtype=f32)
            if want_z_fine:
                out["z_fine"] = torch.empty(n, self.n_fine, device=dev, dtype=f32)
        o = NsrOutputs()
        for k, v in out.items():
            setattr(o, k, v.data_ptr())
        r = NsrRng()
        keep = []
        if rng is not None:
            for k in ("u_coarse", "noise_coarse", "u_fine", "noise_fine"):
                t = rng.get(k) if isinstance(rng, Mapping) else getattr(rng, k, None)
                if t is not None:
                    t = self._f32(t.to(dev))
                    keep.append(t)
                    setattr(r, k, t.data_ptr())
        ws, ws_bytes = self._workspace(n)
        self._check(self.lib.nsr_render(self._h, rays.data_ptr(), n, stride, C.byref(r) if rng is not None else None,
                                        C.byref(o), ws, ws_bytes, self._stream()))
        return out

    def set_debug_flags(self, flags: int) -> None:
        """nsr_debug_set_flags: bit 0 = the weight producer skips its copies (timing experiment, wrong results);
        bit 6 (64) = keep the separate coarse / fine / box-average launches where the one-launch frame kernel would run
        (A/B runs and the bit-equality tests)."""
        self._check(self.lib.nsr_debug_set_flags(self._h, int(flags)))

    def render_frame(self, rays: Optional[torch.Tensor] = None, s: int = 1, pose=None, H: int = 0, W: int = 0,
                     focal: float = 0.0, ndc: bool = False, near: float = 2.0, far: float = 6.0,
                     use_pixel_centers: bool = True, unified_dir: bool = False, want_hr: bool = True,
                     rng: Optional[Mapping[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """One frame through nsr_render_frame: forward() + comp_low_res_output (models/nerf_downX_model.py:316-348) and,
        with `pose` instead of `rays`, the dataset's ray generation too -- ONE kernel launch where the option set allows
        (tensor-core precision, 64 + 64 samples, s in {1, 2, 4}), the same results from separate launches otherwise.
        Returns device tensors: {coarse,fine}_comp_rgbs / _depth / _opacity (HR, when want_hr) and, for s > 1,
        {coarse,fine}_lr_rgb [N/s^2, 3] / _lr_depth [N/s^2]."""
        dev, f32 = self.device, torch.float32
        spec, arr = None, None
        if rays is not None:
            rays = self._f32(rays)
            n, stride = rays.shape
        else:
            c = torch.as_tensor(pose, dtype=torch.float32).reshape(12).cpu()
            arr = (C.c_float * 12)(*c.tolist())
            spec = NsrRayGen()
            spec.struct_size = C.sizeof(NsrRayGen)
            spec.H, spec.W, spec.s, spec.focal, spec.ndc = H, W, s, float(focal), int(ndc)
            spec.near_plane, spec.far_plane = float(near), float(far)
            spec.use_pixel_centers, spec.unified_dir = int(use_pixel_centers), int(unified_dir)
            n, stride = H * W, 8
        if n % (s * s):
            raise NsrError(1, f"{n} rays are not a multiple of s*s = {s * s}")
        names = ("coarse", "fine") if self.n_importance > 0 else ("coarse",)
        out: Dict[str, torch.Tensor] = {}
        o, l = NsrOutputs(), NsrLrOutputs()
        for name in names:
            if want_hr or s == 1:
                out[f"{name}_comp_rgbs"] = torch.empty(n, 3, device=dev, dtype=f32)
                out[f"{name}_depth"] = torch.empty(n, device=dev, dtype=f32)
                out[f"{name}_opacity"] = torch.empty(n, device=dev, dtype=f32)
                for k in ("comp_rgbs", "depth", "opacity"):
                    setattr(o, f"{name}_{k}", out[f"{name}_{k}"].data_ptr())
            if s > 1:
                out[f"{name}_lr_rgb"] = torch.empty(n // (s * s), 3, device=dev, dtype=f32)
                out[f"{name}_lr_depth"] = torch.empty(n // (s * s), device=dev, dtype=f32)
                setattr(l, f"{name}_rgb", out[f"{name}_lr_rgb"].data_ptr())
                setattr(l, f"{name}_depth", out[f"{name}_lr_depth"].data_ptr())
        r = NsrRng()
        keep = []
        if rng is not None:
            for k in ("u_coarse", "noise_coarse", "u_fine", "noise_fine"):
                t = rng.get(k) if isinstance(rng, Mapping) else getattr(rng, k, None)
                if t is not None:
                    t = self._f32(t.to(dev))
                    keep.append(t)
                    setattr(r, k, t.data_ptr())
        need = self.lib.nsr_frame_workspace_bytes(self._h, n, int(rays is None))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        self._check(self.lib.nsr_render_frame(self._h, _ptr(rays), n, stride, arr, C.byref(spec) if spec is not None else None,
                                              s, C.byref(r) if rng is not None else None, C.byref(o), C.byref(l),
                                              self._ws.data_ptr(), self._ws.numel(), self._stream()))
        return out

    def render_pass(self, which: int, rays: torch.Tensor, z_vals: torch.Tensor,
                    noise: Optional[torch.Tensor] = None, want_raw: bool = False) -> Dict[str, torch.Tensor]:
        """One network over caller-supplied z-values: (comp_rgbs, depth, opacity, weights[, raw])."""
        rays, z_vals = self._f32(rays), self._f32(z_vals)
        n, stride = rays.shape
        s = z_vals.shape[1]
        dev, f32 = self.device, torch.float32
        out = {"comp_rgbs": torch.empty(n, 3, device=dev, dtype=f32), "depth": torch.empty(n, device=dev, dtype=f32),
               "opacity": torch.empty(n, device=dev, dtype=f32), "weights": torch.empty(n, s, device=dev, dtype=f32)}
        if want_raw:
            out["raw"] = torch.empty(n, s, 4, device=dev, dtype=f32)
        o = NsrPassOutputs()
        for k, v in out.items():
            setattr(o, k, v.data_ptr())
        nz = self._f32(noise) if noise is not None else None
        ws, ws_bytes = self._workspace(n)
        self._check(self.lib.nsr_render_pass(self._h, which, rays.data_ptr(), n, stride, z_vals.data_ptr(), s,
                                             _ptr(nz), C.byref(o), ws, ws_bytes, self._stream()))
        return out

    def sample_along_rays(self, rays: torch.Tensor, u: Optional[torch.Tensor] = None) -> torch.Tensor:
        rays = self._f32(rays)
        z = torch.empty(rays.shape[0], self.n_coarse, device=self.device, dtype=torch.float32)
        uu = self._f32(u) if u is not None else None
        self._check(self.lib.nsr_sample_coarse(self._h, rays.data_ptr(), rays.shape[0], rays.shape[1], _ptr(uu),
                                               z.data_ptr(), self._stream()))
        return z

    def resample_along_rays(self, z: torch.Tensor, weights: torch.Tensor, u: Optional[torch.Tensor] = None):
        z, weights = self._f32(z), self._f32(weights)
        out = torch.empty(z.shape[0], self.n_fine, device=self.device, dtype=torch.float32)
        uu = self._f32(u) if u is not None else None
        self._check(self.lib.nsr_resample(self._h, z.data_ptr(), weights.data_ptr(), z.shape[0], _ptr(uu),
                                          out.data_ptr(), self._stream()))
        return out

    def posenc(self, x: torch.Tensor, deg: int) -> torch.Tensor:
        x = self._f32(x)
        ch = 6 * deg + (0 if self.cfg.no_xyz else 3)
        out = torch.empty(x.shape[0], ch, device=self.device, dtype=torch.float32)
        self._check(self.lib.nsr_posenc(self._h, x.data_ptr(), x.shape[0], deg, out.data_ptr(), self._stream()))
        return out

    def box_average(self, x: torch.Tensor, s: int) -> torch.Tensor:
        x = self._f32(x)
        x2 = x.reshape(x.shape[0], -1)
        n_lr = x2.shape[0] // (s * s)
        out = torch.empty(n_lr, x2.shape[1], device=self.device, dtype=torch.float32)
        self._check(self.lib.nsr_box_average(self._h, x2.data_ptr(), n_lr, s, x2.shape[1], out.data_ptr(), self._stream()))
        return out

    def lr_metrics(self, hr_rgb: torch.Tensor, target_lr: torch.Tensor, s: int):
        """comp_low_res_output + ColorMSELoss + PSNR (models/nerf_downX_model.py:337-340,357,380):
        returns (lr_rgb [n_lr,3], metrics [2] = (mse, psnr)) as device tensors, no host sync."""
        hr_rgb, target_lr = self._f32(hr_rgb), self._f32(target_lr)
        n_lr = target_lr.shape[0]
        if hr_rgb.shape[0] != n_lr * s * s:
            raise NsrError(1, f"hr_rgb has {hr_rgb.shape[0]} rows, expected {n_lr}*{s}*{s}")
        lr = torch.empty(n_lr, 3, device=self.device, dtype=torch.float32)
        m = torch.empty(2, device=self.device, dtype=torch.float32)
        self._check(self.lib.nsr_lr_metrics(self._h, hr_rgb.data_ptr(), target_lr.data_ptr(), n_lr, s, lr.data_ptr(),
                                            m.data_ptr(), self._stream()))
        return lr, m

    def assemble_frame(self, rgb: torch.Tensor, depth: torch.Tensor, H: int, W: int, s: int, near: float, far: float,
                       gt: Optional[torch.Tensor] = None, want_depth_mat: bool = True):
        """calculate_vis + _save_image conversion (models/nerf_downX_model.py:410-450, utils/visualizer.py:40-60,
        164-176): (uint8 [H, W*(2|3), 3] = [pred | gt | JET depth], fp32 depth matrix [H, W]) on the device."""
        rgb, depth = self._f32(rgb), self._f32(depth)
        if rgb.shape[0] != H * W or depth.numel() != H * W:
            raise NsrError(1, f"expected {H}*{W} rows")
        g = self._f32(gt) if gt is not None else None
        panels = 3 if gt is not None else 2
        out = torch.empty(H, W * panels, 3, dtype=torch.uint8, device=self.device)
        mat = torch.empty(H, W, dtype=torch.float32, device=self.device) if want_depth_mat else None
        self._check(self.lib.nsr_assemble_frame(self._h, rgb.data_ptr(), depth.data_ptr(), _ptr(g), H, W, s, float(near), float(far),
                                                out.data_ptr(), _ptr(mat), self._stream()))
        return out, mat

    def generate_rays(self, c2w, H: int, W: int, focal: float, s: int = 1, ndc: bool = False,
                      near: float = 2.0, far: float = 6.0, use_pixel_centers: bool = True,
                      unified_dir: bool = False) -> torch.Tensor:
        """[H*W, 8] device rays of one pose, LR-pixel-major / sub-pixel-minor rows (nsr_generate_rays; the
        --use_pixel_centers False / --unified_dir variants go through nsr_generate_rays_ex)."""
        c = torch.as_tensor(c2w, dtype=torch.float32).reshape(12).cpu()
        arr = (C.c_float * 12)(*c.tolist())
        rays = torch.empty(H * W, 8, device=self.device, dtype=torch.float32)
        if use_pixel_centers and not unified_dir:
            self._check(self.lib.nsr_generate_rays(self._h, arr, H, W, float(focal), s, int(ndc), float(near), float(far),
                                                   rays.data_ptr(), self._stream()))
        else:
            spec = NsrRayGen()
            spec.struct_size = C.sizeof(NsrRayGen)
            spec.H, spec.W, spec.s, spec.focal, spec.ndc = H, W, s, float(focal), int(ndc)
            spec.near_plane, spec.far_plane = float(near), float(far)
            spec.use_pixel_centers, spec.unified_dir = int(use_pixel_centers), int(unified_dir)
            self._check(self.lib.nsr_generate_rays_ex(self._h, arr, C.byref(spec), rays.data_ptr(), self._stream()))
        return rays

    def render_frame_host(self, rays_host: torch.Tensor, s: int = 1):
        """HOST rays [N, 8|11] (CPU tensor) -> (rgb [N/s^2,3], depth [N/s^2]) CPU tensors.
        Copies, render and box average are pipelined inside the library."""
        rays_host = rays_host.detach().to("cpu", torch.float32).contiguous()
        n, stride = rays_host.shape
        n_out = n // (s * s)
        rgb = torch.empty(n_out, 3, dtype=torch.float32)
        depth = torch.empty(n_out, dtype=torch.float32)
        self._check(self.lib.nsr_render_host(self._h, rays_host.data_ptr(), n, stride, s, rgb.data_ptr(), depth.data_ptr()))
        return rgb, depth

    def render_pose_host(self, c2w, H: int, W: int, focal: float, s: int = 1, ndc: bool = False,
                          near: float = 2.0, far: float = 6.0):
        """One camera pose -> (rgb [H*W/s^2,3], depth [H*W/s^2]) CPU tensors; rays are generated on the
        device (48 bytes of pose go up instead of 32 bytes per ray)."""
        c = torch.as_tensor(c2w, dtype=torch.float32).reshape(12).cpu()
        arr = (C.c_float * 12)(*c.tolist())
        n_out = (H // s) * (W // s)
        rgb = torch.empty(n_out, 3, dtype=torch.float32)
        depth = torch.empty(n_out, dtype=torch.float32)
        self._check(self.lib.nsr_render_pose_host(self._h, arr, H, W, float(focal), s, int(ndc), float(near), float(far),
                                                  rgb.data_ptr(), depth.data_ptr()))
        return rgb, depth

    def render_test_pose(self, c2w, H: int, W: int, focal: float, s: int = 2, ndc: bool = False,
                          near: float = 2.0, far: float = 6.0) -> Dict[str, torch.Tensor]:
        """One iteration of the reference's test sweep for a camera pose, entirely on the device:
        dataset __getitem__ (rays for the pose, data/*_downX_dataset.py test branch) -> forward -> comp_low_res_output
        -> calculate_vis(with_gt=False) (models/nerf_downX_model.py:316-353,418-450).  Returns device tensors:
        {coarse,fine}_pred_ori  uint8 [H, 2W, 3]      HR [pred | depth] frames
        {coarse,fine}_pred      uint8 [H/s, 2W/s, 3]  LR (box-averaged) frames
        {coarse,fine}_depth_mat_ori [H, W], {coarse,fine}_depth_mat [H/s, W/s]   fp32 depth matrices (the *.npz payloads)"""
        nr, fr = (0.0, 1.0) if ndc else (near, far)              # self.near / self.far = rays[0, 6:8]
        # rays, both passes and the box averages: one launch on the tensor-core path (nsr_render_frame)
        out = self.render_frame(None, s, pose=c2w, H=H, W=W, focal=focal, ndc=ndc, near=near, far=far)
        res: Dict[str, torch.Tensor] = {}
        for name in ("coarse", "fine") if self.n_importance > 0 else ("coarse",):
            rgb, depth = out[f"{name}_comp_rgbs"], out[f"{name}_depth"]
            res[f"{name}_pred_ori"], res[f"{name}_depth_mat_ori"] = self.assemble_frame(rgb, depth, H, W, s, nr, fr)
            lr_rgb, lr_depth = (out[f"{name}_lr_rgb"], out[f"{name}_lr_depth"]) if s > 1 else (rgb, depth)
            res[f"{name}_pred"], res[f"{name}_depth_mat"] = self.assemble_frame(lr_rgb, lr_depth, H // s, W // s, 1, nr, fr)
        return res

    def render_path(self, poses, H: int, W: int, focal: float, s: int = 2, ndc: bool = False, near: float = 2.0,
                     far: float = 6.0, keys: Sequence[str] = ("fine_pred", "fine_pred_ori", "fine_depth_mat_ori")):
        """Test sweep over a pose path (nerf_sr_b200.paths): yields one dict of HOST (pinned) tensors per pose.
        48 bytes go up per frame; frame k's device-to-host copies overlap frame k+1's render (two pinned slots)."""
        copy_stream = torch.cuda.Stream(self.device)
        slots, events, pending = [None, None], [None, None], None
        for k, c2w in enumerate(poses):
            res = self.render_test_pose(c2w, H, W, focal, s, ndc, near, far)
            slot = k & 1
            if slots[slot] is None:
                slots[slot] = {n: torch.empty(res[n].shape, dtype=res[n].dtype).pin_memory() for n in keys}
            done = torch.cuda.Event()
            copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(copy_stream):
                for n in keys:
                    res[n].record_stream(copy_stream)
                    slots[slot][n].copy_(res[n], non_blocking=True)
                done.record(copy_stream)
            events[slot] = done
            if pending is not None:
                events[pending].synchronize()
                yield slots[pending]
            pending = slot
        if pending is not None:
            events[pending].synchronize()
            yield slots[pending]


class _LazyNearFar:
    """``self.near`` / ``self.far`` of the reference model (models/nerf_downX_model.py:284: ``near[0].cpu().numpy()``,
    shape-(1,) numpy arrays read by depth2im :422 and the depth-variance loss :351) without the reference's blocking
    device read per forward_rays call: the 8 bytes are copied to pinned memory on the render stream and the event is
    waited on only when somebody reads the attribute."""

    def __init__(self):
        self._pinned = None
        self._event = None
        self._override = {}

    def capture(self, rays: torch.Tensor):
        self._override.clear()
        if rays.device.type != "cuda":
            self._pinned, self._event = rays[0, 6:8].detach().float().clone(), None
            return
        if self._pinned is None or not self._pinned.is_pinned():
            self._pinned = torch.empty(2, dtype=torch.float32).pin_memory()
        if self._event is not None:
            self._event.synchronize()          # the previous copy must have landed before the slot is reused
        self._pinned.copy_(rays[0, 6:8].detach(), non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record(torch.cuda.current_stream(rays.device))

    def get(self, i: int):
        if i in self._override:
            return self._override[i]
        if self._pinned is None:
            raise AttributeError("near / far are set by the first forward_rays call")
        if self._event is not None:
            self._event.synchronize()
        return self._pinned[i:i + 1].numpy().copy()

    def set(self, i: int, v):
        self._override[i] = v


def _install_lazy_near_far(model) -> _LazyNearFar:
    """Give ``model`` lazily evaluated ``near`` / ``far`` attributes (a per-instance subclass with two properties; the
    class name is kept, assignments by other code -- the reference's own forward_rays on the fallback path -- still work)."""
    lazy = _LazyNearFar()
    cls = type(model)
    sub = type(cls.__name__, (cls,), {
        "near": property(lambda self: lazy.get(0), lambda self, v: lazy.set(0, v)),
        "far": property(lambda self: lazy.get(1), lambda self, v: lazy.set(1, v)),
        "__module__": cls.__module__})
    for k in ("near", "far"):
        model.__dict__.pop(k, None)
    model.__class__ = sub
    return lazy


def _ddp_group(net):
    """The process group of a DistributedDataParallel wrapper (models/networks.py:72-86), else None."""
    try:
        from torch.nn.parallel import DistributedDataParallel as DDP
    except Exception:                           # pragma: no cover
        return None
    if isinstance(net, DDP) and torch.distributed.is_initialized() and torch.distributed.get_world_size(net.process_group) > 1:
        return net.process_group
    return None


def patch_model(model, precision: str = "bf16x3", whole_frame: bool = True):
    """Rebind ``forward_rays`` of a reference NeRFDownXModel / NeRFModel instance to the CUDA path
    (the one-line hook of INTEGRATION.md).

    * ``self.near`` / ``self.far`` (depth2im, models/nerf_downX_model.py:422; depth-variance loss :351) keep the reference's
      value and type but are read lazily (``_LazyNearFar``): no device sync per call, where the reference has one per
      4096-ray chunk (:284).
    * With grad enabled the outputs are autograd-connected to the parameters of netCoarse / netFine through
      training.RenderFunction (CUDA backward), so the reference's ``loss_tot.backward()`` / ``optimizer.step()`` run
      unchanged.  When the nets are wrapped in DistributedDataParallel (``--accelerator ddp``, models/networks.py:72-86) the
      backward averages the two flat gradient buffers over the DDP process group -- what DDP's reducer would have done had
      its ``forward`` been called -- so every rank steps with the same gradients.
    * Option sets the library refuses (``nsr_create`` -> NSR_ERR_UNSUPPORTED) leave the model untouched (one warning);
      option sets only the backward does not cover (precisions other than bf16x3, sample counts other than 64 + 64, --no_dir) keep the
      reference path in train mode, chunked by the ORIGINAL ``opt.ray_chunk``.
    * whole_frame: raise ``opt.ray_chunk`` so that the reference's ``chunk_batch(self.forward_rays, opt.ray_chunk, rays)``
      (models/nerf_downX_model.py:318, utils/utils.py:130-152) hands a whole frame to one call -- the chunking only exists to
      bound the [P, 90] intermediates of the PyTorch path, which this path never materialises; the stitched result is the same."""
    import types
    import warnings
    vo = 8 if type(model).__name__ == "NeRFModel" else 3
    try:
        renderer = Renderer(model.opt, device=model.device, precision=precision, viewdir_offset=vo)
    except NsrError as e:
        if e.code != 2:
            raise
        warnings.warn(f"nerf_sr_b200.patch_model: option set not supported by the CUDA path ({e}); "
                      "the reference forward_rays stays in place", RuntimeWarning, stacklevel=2)
        model._nsr_renderer = None
        return model
    reference_chunk = int(getattr(model.opt, "ray_chunk", 4096))
    if whole_frame and hasattr(model.opt, "ray_chunk"):
        model.opt.ray_chunk = max(reference_chunk, 1 << 30)
    reference_forward_rays = model.forward_rays
    lazy = _install_lazy_near_far(model)

    train_capable = (precision == "bf16x3" and (renderer.n_coarse, renderer.n_importance) == (64, 64) and not renderer.cfg.no_dir
                     and renderer.cfg.W == 256)

    def reference_path(rays):
        # the PyTorch path needs its chunking back (utils/utils.py:130-152) when whole_frame lifted opt.ray_chunk
        if rays.shape[0] <= reference_chunk:
            return reference_forward_rays(rays)
        parts = [reference_forward_rays(rays[i:i + reference_chunk]) for i in range(0, rays.shape[0], reference_chunk)]
        return {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}

    def forward_rays(self, rays):
        grad_mode = torch.is_grad_enabled() and any(
            p.requires_grad for net in (self.netCoarse, self.netFine) for p in net.parameters())
        if grad_mode and not train_capable:
            return reference_path(rays)
        if not grad_mode:
            renderer.sync_from_modules(self.netCoarse, self.netFine)
        lazy.capture(rays)                         # 8 bytes, asynchronous (the reference: a blocking read per chunk, :284)
        rng = None
        if self.randomized:
            n = rays.shape[0]
            o = self.opt
            rng = {"u_coarse": torch.rand(n, o.N_coarse, device=rays.device)}
            if o.noise_std > 0:
                rng["noise_coarse"] = torch.randn(n, o.N_coarse, device=rays.device)
            if o.N_importance > 0:
                rng["u_fine"] = torch.rand(n, o.N_importance, device=rays.device)
                if o.noise_std > 0:
                    rng["noise_fine"] = torch.randn(n, o.N_coarse + o.N_importance, device=rays.device)
        if grad_mode:
            from . import training as T
            pcs, pfs = T.module_params_in_order(self.netCoarse), T.module_params_in_order(self.netFine)
            group = _ddp_group(self.netCoarse)
            if group is None:
                group = _ddp_group(self.netFine)
            outs = T.RenderFunction.apply(renderer, rays, rng, len(pcs), group, *pcs, *pfs)
            renderer._param_versions = [None, None]          # the images now hold the training weights
            return dict(zip(T.OUT_KEYS, outs))
        return renderer.forward_rays(rays, rng)

    model.forward_rays = types.MethodType(forward_rays, model)
    model._nsr_renderer = renderer
    return model
