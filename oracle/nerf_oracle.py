"""CPU oracle for the NeRF-SR volumetric-render hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``nerf_sr_b200/`` imports this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may use it, and only as the checker
or as the timed CPU arm -- never as a product fallback.

This is a from-scratch functional restatement (plain ``torch`` CPU ops, fp32)
of the reference algorithm; each function cites the reference ``file:line`` it
follows (paths relative to the cwchenwang/NeRF-SR tree).  The reference keeps
this logic inside ``nn.Module``/model classes driven by an argparse ``opt``;
here it is stateless functions over a parameter dict that uses the reference's
``state_dict`` key names, so the same weights feed the reference, the oracle
and the CUDA library.

Parity pin: the reference ships no golden vectors, so the
oracle is pinned against the reference ITSELF, imported in the build container
through ``oracle/ref_shim.py`` -- ``oracle/make_golden.py`` asserts bit-equal
outputs (same ATen CPU kernels, same op order) and commits the resulting
fixtures under ``tests/golden/``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor

# input generators / option record live in the package (they are not checker logic); re-exported here
# so tests can keep using a single namespace
from nerf_sr_b200.synthetic import (RenderConfig, encoded_channels, make_mlp_params,  # noqa: E402,F401
                                    mlp_param_shapes, synthetic_rays)


# --------------------------------------------------------------------------
# configuration (the option surface the path reads)
# --------------------------------------------------------------------------
# --------------------------------------------------------------------------
# a5: positional encoding   (models/embedding.py:21-26, 39-42, 44-63)
# --------------------------------------------------------------------------
def frequency_bands(n_freqs: int, no_logscale: bool = False) -> Tensor:
    """models/embedding.py:39-42."""
    if no_logscale:
        return torch.linspace(1, 2 ** (n_freqs - 1), n_freqs)
    return 2 ** torch.linspace(0, n_freqs - 1, n_freqs)


def posenc(x: Tensor, n_freqs: int, no_xyz: bool = False,
           no_logscale: bool = False) -> Tensor:
    """[B,C] -> [B, C*(2L)+C]; channel order x, sin f0, cos f0, sin f1, ...
    (models/embedding.py:57-63)."""
    parts: List[Tensor] = [] if no_xyz else [x]
    for f in frequency_bands(n_freqs, no_logscale).to(device=x.device, dtype=x.dtype):
        parts.append(torch.sin(f * x))
        parts.append(torch.cos(f * x))
    return torch.cat(parts, -1)


# --------------------------------------------------------------------------
# a8: the MLP   (models/networks.py:149-180 layers, 199-224 forward)
# --------------------------------------------------------------------------
def _linear(x: Tensor, p: Dict[str, Tensor], name: str) -> Tensor:
    return torch.nn.functional.linear(x, p[name + ".weight"], p[name + ".bias"])


def mlp_forward(p: Dict[str, Tensor], x: Tensor, cfg: RenderConfig, acts: Optional[list] = None,
                relu_masks: Optional[Sequence[Tensor]] = None) -> Tensor:
    """[P, ch_pos+ch_dir] -> [P,4] = (rgb after colour activation, raw sigma).
    models/networks.py:199-224.  ``acts`` (a list) receives the intermediate tensors
    [h_1, ..., h_D, feat, dir_act] for the training-stash tests.
    ``relu_masks`` (test harness only; None = the reference's path, untouched): D + 1 boolean tensors [P, W] / [P, W/2]
    that REPLACE the sign test of the D trunk ReLUs and of the dir layer's ReLU (y = pre * mask) -- "mask teacher
    forcing": the gradient then flows through exactly the units another implementation's forward kept active, so
    a gradient comparison is not polluted by ReLU decisions that flip on pre-activations of ~1e-5."""
    ch_pos, ch_dir = cfg.ch_pos, cfg.ch_dir
    in_xyz, in_dir = torch.split(x, [ch_pos, ch_dir], dim=-1)      # :199
    h = in_xyz
    act = (lambda pre, i: torch.relu(pre)) if relu_masks is None else (lambda pre, i: pre * relu_masks[i].to(pre.dtype))
    for i in range(cfg.D):                                           # :202-205
        if i in cfg.skips:
            h = torch.cat([in_xyz, h], -1)                           # :204
        h = act(_linear(h, p, f"xyz_encoding_{i+1}.0"), i)
        if acts is not None:
            acts.append(h)
    sigma = _linear(h, p, "sigma")                                   # :207
    feat = _linear(h, p, "xyz_encoding_final")                       # :211 (no act)
    d_in = feat if cfg.no_dir else torch.cat([feat, in_dir], -1)     # :213-216
    d = act(_linear(d_in, p, "dir_encoding.0"), cfg.D)               # :221
    if acts is not None:
        acts += [feat, d]
    rgb = _linear(d, p, "rgb.0")                                     # :222
    if cfg.color_activation == "sigmoid":
        rgb = torch.sigmoid(rgb)
    return torch.cat([rgb, sigma], -1)                               # :224


# --------------------------------------------------------------------------
# a6: sampling   (models/utils.py:5-14, 17-44)
# --------------------------------------------------------------------------
def cast_rays(o: Tensor, d: Tensor, z: Tensor) -> Tensor:
    """o + z*d, multiply then add (models/utils.py:14)."""
    return o[..., None, :] + z[..., None] * d[..., None, :]


def sample_along_rays(o: Tensor, d: Tensor, near: Tensor, far: Tensor, n: int,
                      lindisp: bool, u: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """near/far are [N,1].  ``u`` ([N,n] uniform draws) stands in for the
    reference's ``torch.rand_like`` (models/utils.py:41); ``None`` means the
    deterministic eval path."""
    t = torch.linspace(0, 1, n, device=o.device)                     # :31
    if lindisp:
        z = 1. / (1. / near * (1 - t) + 1. / far * t)                # :33
    else:
        z = near * (1 - t) + far * t                                 # :35
    if u is not None:                                                # :37-41
        mids = 0.5 * (z[:, :-1] + z[:, 1:])
        upper = torch.cat([mids, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mids], -1)
        z = lower + u * (upper - lower)
    return z, cast_rays(o, d, z)


# --------------------------------------------------------------------------
# a10: alpha compositing   (models/rendering.py:89-111)
# --------------------------------------------------------------------------
def composite(rgb: Tensor, sigma: Tensor, z: Tensor, white_bkgd: bool,
              sigma_activation: str = "relu", sigma_mask: Optional[Tensor] = None):
    """``sigma_mask`` (test harness only, None = the reference's path): replaces the sign test of relu(sigma) the same way
    mlp_forward's ``relu_masks`` do.  The LAST sample has delta = 1e10, so alpha_last jumps between 0 and 1 with the sign of
    sigma_last: the one genuinely discontinuous decision of the compositing stage."""
    eps = 1e-10
    deltas = z[:, 1:] - z[:, :-1]                                    # :90
    deltas = torch.cat([deltas, 1e10 * torch.ones_like(deltas[:, :1])], -1)
    if sigma_activation == "relu" and sigma_mask is not None:
        act = sigma * sigma_mask.to(sigma.dtype)
    elif sigma_activation == "relu":                                 # :70-73
        act = torch.relu(sigma)
    else:
        act = torch.log(1 + torch.exp(sigma - 1))
    alpha = 1 - torch.exp(-deltas * act)                             # :98
    trans = torch.cat([torch.ones_like(alpha[:, :1]),                # :99-102
                       torch.cumprod(1 - alpha[:, :-1] + eps, dim=-1)], -1)
    weights = alpha * trans                                          # :103
    comp = (weights[..., None] * rgb).sum(dim=-2)                    # :104
    depth = (weights * z).sum(dim=-1)                                # :105
    opacity = weights.sum(dim=-1)                                    # :106
    if white_bkgd:
        comp = comp + (1 - opacity[..., None])                       # :108-109
    return comp, depth, opacity, weights


# --------------------------------------------------------------------------
# a11: inverse-CDF resampling + sort-merge   (models/utils.py:47-95)
# --------------------------------------------------------------------------
def resample_along_rays(o: Tensor, d: Tensor, z: Tensor, weights: Tensor,
                        n: int, u: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """``u`` ([N,n]) replaces ``torch.rand`` (models/utils.py:73); None ->
    linspace(0,1,n) (eval)."""
    eps = 1e-5
    bins = 0.5 * (z[:, :-1] + z[:, 1:])                              # :63
    w = weights[:, 1:-1]                                             # :64
    n_rays, n_w = w.shape
    w = w + eps                                                      # :67
    pdf = w / w.sum(dim=-1, keepdim=True)                            # :68
    cdf = torch.cumsum(pdf, -1)                                      # :69
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)         # :70
    if u is None:
        u = torch.linspace(0, 1, n, device=z.device).expand(n_rays, n)   # :75-76
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)                    # :79
    below = torch.clamp_min(inds - 1, 0)                             # :80
    above = torch.clamp_max(inds, n_w)                               # :81
    pair = torch.stack([below, above], -1).view(n_rays, -1)          # :83
    cdf_g = torch.gather(cdf, 1, pair).view(n_rays, -1, 2)           # :84
    bins_g = torch.gather(bins, 1, pair).view(n_rays, -1, 2)         # :85
    denom = cdf_g[..., 1] - cdf_g[..., 0]                            # :87
    denom[denom < eps] = 1                                           # :88
    z_new = bins_g[..., 0] + (u - cdf_g[..., 0]) / denom * (bins_g[..., 1] - bins_g[..., 0])
    z_all = torch.sort(torch.cat([z, z_new], -1), -1)[0]             # :93
    return z_all, cast_rays(o, d, z_all)


# --------------------------------------------------------------------------
# a7 + a9: one network pass over sampled points
# (models/nerf_downX_model.py:260-278, models/utils.py:199-212)
# --------------------------------------------------------------------------
def render_pass(p: Dict[str, Tensor], xyz: Tensor, dir_enc: Tensor, z: Tensor,
                cfg: RenderConfig, noise: Optional[Tensor], relu_masks: Optional[Sequence[Tensor]] = None):
    n_rays, n_s = xyz.shape[:2]
    enc = posenc(xyz.reshape(-1, 3), cfg.deg_pos, cfg.no_xyz, cfg.no_logscale)
    d = dir_enc.repeat_interleave(n_s, dim=0)                        # ray-major, :266
    sigma_mask = None
    if relu_masks is not None and len(relu_masks) > cfg.D + 1:       # optional last entry: [N, S] mask of relu(sigma)
        sigma_mask, relu_masks = relu_masks[cfg.D + 1], relu_masks[: cfg.D + 1]
    raw = mlp_forward(p, torch.cat([enc, d], -1), cfg, relu_masks=relu_masks).view(n_rays, n_s, 4)
    rgb, sigma = raw[..., :3], raw[..., 3]
    if cfg.gamma_correct:                                            # :271-276
        rgb = torch.pow(rgb, 1 / 2.2)
    if noise is not None and cfg.noise_std > 0:                      # utils.py:209-210
        sigma = sigma + noise * cfg.noise_std
    return composite(rgb, sigma, z, cfg.white_bkgd, cfg.sigma_activation, sigma_mask) + (raw,)


@dataclass
class RenderRng:
    """Explicit random inputs replacing the reference's in-line torch RNG
    draws, in the reference's draw order:
    u_coarse [N,Nc] (utils.py:41), noise_coarse [N,Nc] (utils.py:210),
    u_fine [N,Ni] (utils.py:73), noise_fine [N,Nc+Ni]."""
    u_coarse: Optional[Tensor] = None
    noise_coarse: Optional[Tensor] = None
    u_fine: Optional[Tensor] = None
    noise_fine: Optional[Tensor] = None

    @staticmethod
    def draw(n_rays: int, cfg: RenderConfig, generator: torch.Generator) -> "RenderRng":
        r = RenderRng()
        r.u_coarse = torch.rand(n_rays, cfg.N_coarse, generator=generator)
        if cfg.noise_std > 0:
            r.noise_coarse = torch.randn(n_rays, cfg.N_coarse, generator=generator)
        if cfg.N_importance > 0:
            r.u_fine = torch.rand(n_rays, cfg.N_importance, generator=generator)
            if cfg.noise_std > 0:
                r.noise_fine = torch.randn(n_rays, cfg.N_coarse + cfg.N_importance,
                                           generator=generator)
        return r


# --------------------------------------------------------------------------
# a12: forward_rays   (models/nerf_downX_model.py:280-313)
# --------------------------------------------------------------------------
def forward_rays(p_coarse: Dict[str, Tensor], p_fine: Dict[str, Tensor],
                 rays: Tensor, cfg: RenderConfig,
                 rng: Optional[RenderRng] = None,
                 z_fine_override: Optional[Tensor] = None,
                 extras: Optional[dict] = None,
                 relu_masks: Optional[Tuple[Optional[Sequence[Tensor]], Optional[Sequence[Tensor]]]] = None) -> Dict[str, Tensor]:
    """rays [N, 8|11] = (o3, d3, near, far[, viewdir3]) -> the reference's
    8-key dict.  ``rng=None`` is eval mode.  ``z_fine_override`` teacher-forces
    the fine pass (parity protocol ii).  ``extras`` (a
    dict) receives intermediate tensors (z_coarse, z_fine, raw_coarse, raw_fine)."""
    o, d = rays[:, 0:3], rays[:, 3:6]                                # :282
    near, far = rays[:, 6:7], rays[:, 7:8]                           # :283
    vo = cfg.viewdir_offset
    dir_enc = posenc(rays[:, vo:vo + 3], cfg.deg_dir, cfg.no_xyz, cfg.no_logscale)  # :286
    rng = rng or RenderRng()
    z, xyz = sample_along_rays(o, d, near, far, cfg.N_coarse, cfg.lindisp, rng.u_coarse)
    rm = relu_masks or (None, None)        # (coarse, fine) mask teacher forcing, see mlp_forward
    c_rgb, c_depth, c_opa, c_w, raw_c = render_pass(p_coarse, xyz, dir_enc, z, cfg,
                                                    rng.noise_coarse, rm[0])
    out = {"coarse_comp_rgbs": c_rgb, "coarse_depth": c_depth,
           "coarse_opacity": c_opa, "coarse_weights": c_w}           # :293-298
    if extras is not None:
        extras["z_coarse"], extras["raw_coarse"] = z, raw_c
    if cfg.N_importance > 0:                                         # :300-311
        if z_fine_override is None:
            z_f, xyz_f = resample_along_rays(o, d, z, c_w.detach(), cfg.N_importance,
                                             rng.u_fine)
        else:
            z_f, xyz_f = z_fine_override, cast_rays(o, d, z_fine_override)
        f_rgb, f_depth, f_opa, f_w, raw_f = render_pass(p_fine, xyz_f, dir_enc, z_f, cfg,
                                                        rng.noise_fine, rm[1])
        out.update({"fine_comp_rgbs": f_rgb, "fine_depth": f_depth,
                    "fine_opacity": f_opa, "fine_weights": f_w})
        if extras is not None:
            extras["z_fine"], extras["raw_fine"] = z_f, raw_f
    return out


def chunked_forward(p_coarse, p_fine, rays: Tensor, cfg: RenderConfig,
                    ray_chunk: int = 4096) -> Dict[str, Tensor]:
    """a13: utils/utils.py:130-152 + models/nerf_downX_model.py:316-319
    (eval mode only)."""
    acc: Dict[str, List[Tensor]] = {}
    for i in range(0, rays.shape[0], ray_chunk):
        for k, v in forward_rays(p_coarse, p_fine, rays[i:i + ray_chunk], cfg).items():
            acc.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in acc.items()}


# --------------------------------------------------------------------------
# a14: s x s box average   (models/nerf_downX_model.py:337-348)
# --------------------------------------------------------------------------
def box_average(x: Tensor, s: int) -> Tensor:
    """[N, C] or [N] (N = n_lr * s*s, sub-pixels contiguous) -> [n_lr, C|1]."""
    n_lr = x.shape[0] // (s * s)
    return torch.mean(torch.reshape(x, (n_lr, s * s, -1)), dim=1)


def lr_metrics(hr_rgb: Tensor, target_lr: Tensor, s: int):
    """Scope row f-2: box average (models/nerf_downX_model.py:337-340), ColorMSELoss
    (models/criterions.py:7-15, nn.MSELoss mean) and PSNR (models/criterions.py:27-36)."""
    lr = box_average(hr_rgb, s)
    mse = torch.nn.functional.mse_loss(lr, target_lr, reduction="mean")
    psnr = -10 * torch.log10(torch.mean((lr - target_lr) ** 2))
    return lr, mse, psnr


# --------------------------------------------------------------------------
# a1-a4: ray generation   (models/utils.py:98-196, data/*_downX_dataset.py)
# --------------------------------------------------------------------------
def ray_directions(H: int, W: int, focal: float, use_pixel_centers: bool = True) -> Tensor:
    """models/utils.py:98-126."""
    c = 0.5 if use_pixel_centers else 0
    i, j = np.meshgrid(np.arange(W, dtype=np.float32) + c,
                       np.arange(H, dtype=np.float32) + c, indexing="xy")
    i, j = torch.from_numpy(i), torch.from_numpy(j)
    return torch.stack([(i - W / 2) / focal, -(j - H / 2) / focal, -torch.ones_like(i)], -1)


def rays_from_pose(directions: Tensor, c2w: Tensor) -> Tuple[Tensor, Tensor]:
    """models/utils.py:129-152: rotate, L2-normalise, broadcast origin."""
    d = directions @ c2w[:, :3].T
    d = d / torch.norm(d, dim=-1, keepdim=True)
    o = c2w[:, 3].expand(d.shape)
    return o.reshape(-1, 3), d.reshape(-1, 3)


def ndc_rays(H: int, W: int, focal: float, near: float, o: Tensor, d: Tensor):
    """models/utils.py:155-196."""
    t = -(near + o[..., 2]) / d[..., 2]
    o = o + t[..., None] * d
    ox_oz = o[..., 0] / o[..., 2]
    oy_oz = o[..., 1] / o[..., 2]
    o0 = -1. / (W / (2. * focal)) * ox_oz
    o1 = -1. / (H / (2. * focal)) * oy_oz
    o2 = 1. + 2. * near / o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (d[..., 0] / d[..., 2] - ox_oz)
    d1 = -1. / (H / (2. * focal)) * (d[..., 1] / d[..., 2] - oy_oz)
    d2 = 1 - o2
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def build_frame_rays(c2w: Tensor, H: int, W: int, focal: float, s: int,
                     near: float, far: float, ndc: bool = False, use_pixel_centers: bool = True,
                     unified_dir: bool = False) -> Tensor:
    """HR raster of rays for one pose, grouped LR-pixel-major / sub-pixel-minor.

    Follows the test branch of data/blender_downX_dataset.py:207-215 (ndc=False:
    cat(o, d, near, far)) and data/llff_downX_dataset.py:473-490 (ndc=True:
    get_ndc_rays at near=1.0, then near=0 far=1), then the einops grouping
    '(h s1) (w s2) c -> (h w) (s1 s2) c' and the flatten of
    models/nerf_downX_model.py:247.  Returns [H*W, 8].
    ``unified_dir`` (data/llff_downX_dataset.py:273-277): directions from the (H/s, W/s) raster with focal // s,
    each repeated over its s x s sub-pixels ('h w c -> (h s1) (w s2) c')."""
    if unified_dir:
        dirs = ray_directions(H // s, W // s, focal // s, use_pixel_centers)
        dirs = dirs.repeat_interleave(s, dim=0).repeat_interleave(s, dim=1)
    else:
        dirs = ray_directions(H, W, focal, use_pixel_centers)
    o, d = rays_from_pose(dirs, c2w)
    if ndc:
        o, d = ndc_rays(H, W, focal, 1.0, o, d)
        near_t, far_t = torch.zeros_like(o[:, :1]), torch.ones_like(o[:, :1])
    else:
        near_t = near * torch.ones_like(o[:, :1])
        far_t = far * torch.ones_like(o[:, :1])
    rays = torch.cat([o, d, near_t, far_t], 1).view(H, W, 8)
    h, w = H // s, W // s
    rays = rays.view(h, s, w, s, 8).permute(0, 2, 1, 3, 4).reshape(h * w * s * s, 8)
    return rays.contiguous()


# --------------------------------------------------------------------------
# synthetic fixtures   
# --------------------------------------------------------------------------
def tolerance_violations(a: Tensor, b: Tensor, rtol: float = 1e-3, atol: float = 1e-4):
    """|a-b| <= atol + rtol*|b|.  Returns
    (max_abs_err, violation_fraction)."""
    a, b = a.double(), b.double()
    err = (a - b).abs()
    bad = err > (atol + rtol * b.abs())
    return float(err.max()) if err.numel() else 0.0, float(bad.double().mean()) if err.numel() else 0.0
