"""Synthetic inputs and the option record shared by bench.py, the tools and the tests.

Not part of the oracle: this module only GENERATES inputs (seeded rays, kaiming-initialised weights in
the reference's state_dict layout, and a dataclass mirroring the subset of the reference's ``opt`` the
render path reads).  It contains no rendering arithmetic.  numpy's PCG64 is used so that the same
(seed, config) reproduces the same fixture on every box without storing 4.8 MB of weights."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np
import torch

Tensor = torch.Tensor


def encoded_channels(in_ch: int, n_freqs: int, no_xyz: bool = False) -> int:
    """models/embedding.py:21-26."""
    return in_ch * 2 * n_freqs + (0 if no_xyz else in_ch)



@dataclass
class RenderConfig:
    """Subset of the reference's ``opt`` that the hot path reads.

    Defaults follow models/nerf_model.py:46-67, models/networks.py:124-128,
    models/embedding.py:17-18, models/nerf_downX_model.py:125.
    """
    D: int = 8
    W: int = 256
    skips: Tuple[int, ...] = (4,)
    deg_pos: int = 10
    deg_dir: int = 4
    no_xyz: bool = False
    no_logscale: bool = False
    no_dir: bool = False
    N_coarse: int = 64
    N_importance: int = 64
    lindisp: bool = False
    noise_std: float = 0.0
    white_bkgd: bool = False
    sigma_activation: str = "relu"      # relu | softplus
    color_activation: str = "sigmoid"   # sigmoid | none
    gamma_correct: bool = False
    downscale: int = 2
    viewdir_offset: int = 3             # downX model: rays[:,3:6]; vanilla: 8

    @property
    def ch_pos(self) -> int:
        return encoded_channels(3, self.deg_pos, self.no_xyz)

    @property
    def ch_dir(self) -> int:
        return encoded_channels(3, self.deg_dir, self.no_xyz)


def mlp_param_shapes(cfg: RenderConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict names and shapes in registration order
    (models/networks.py:149-180)."""
    out: List[Tuple[str, Tuple[int, ...]]] = []
    for i in range(cfg.D):
        if i == 0:
            k = cfg.ch_pos
        elif i in cfg.skips:
            k = cfg.W + cfg.ch_pos
        else:
            k = cfg.W
        out += [(f"xyz_encoding_{i+1}.0.weight", (cfg.W, k)),
                (f"xyz_encoding_{i+1}.0.bias", (cfg.W,))]
    out += [("xyz_encoding_final.weight", (cfg.W, cfg.W)),
            ("xyz_encoding_final.bias", (cfg.W,))]
    kd = cfg.W if cfg.no_dir else cfg.W + cfg.ch_dir
    out += [("dir_encoding.0.weight", (cfg.W // 2, kd)),
            ("dir_encoding.0.bias", (cfg.W // 2,))]
    out += [("sigma.weight", (1, cfg.W)), ("sigma.bias", (1,))]
    out += [("rgb.0.weight", (3, cfg.W // 2)), ("rgb.0.bias", (3,))]
    return out


def make_mlp_params(cfg: RenderConfig, seed: int, sigma_bias: float = 0.0,
                    bias_std: float = 0.0) -> Dict[str, Tensor]:
    """Synthetic weights: kaiming-normal(fan_in, relu) like the reference's
    default init (models/networks.py:31-38: std = sqrt(2/fan_in), bias 0), but
    drawn from numpy's PCG64 so fixtures are reproducible without storing
    4.8 MB of weights.  ``bias_std``/``sigma_bias`` give "trained-like"
    variants with non-zero biases."""
    rng = np.random.default_rng(seed)
    params: Dict[str, Tensor] = {}
    for name, shape in mlp_param_shapes(cfg):
        if name.endswith("weight"):
            std = (2.0 / shape[1]) ** 0.5
            arr = rng.standard_normal(shape, dtype=np.float32) * np.float32(std)
        else:
            arr = rng.standard_normal(shape, dtype=np.float32) * np.float32(bias_std)
            if name == "sigma.bias":
                arr = arr + np.float32(sigma_bias)
        params[name] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
    return params


def synthetic_rays(n: int, seed: int, kind: str = "blender") -> Tensor:
    """'blender': camera on a radius-4 sphere looking at a jittered target,
    near=2 far=6; 'llff': NDC-like, o=(x,y,-1), d_z=2, near=0 far=1."""
    rng = np.random.default_rng(seed)
    if kind == "blender":
        c = rng.standard_normal((n, 3)).astype(np.float32)
        c = 4.0 * c / np.linalg.norm(c, axis=1, keepdims=True)
        tgt = (0.6 * rng.standard_normal((n, 3))).astype(np.float32)
        d = tgt - c
        d = d / np.linalg.norm(d, axis=1, keepdims=True)
        near = np.full((n, 1), 2.0, np.float32)
        far = np.full((n, 1), 6.0, np.float32)
        rays = np.concatenate([c, d, near, far], 1)
    elif kind == "llff":
        xy = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        o = np.concatenate([xy, -np.ones((n, 1), np.float32)], 1)
        dxy = (0.3 * rng.standard_normal((n, 2))).astype(np.float32)
        d = np.concatenate([dxy, 2 * np.ones((n, 1), np.float32)], 1)
        rays = np.concatenate([o, d, np.zeros((n, 1), np.float32),
                               np.ones((n, 1), np.float32)], 1)
    else:
        raise ValueError(kind)
    return torch.from_numpy(np.ascontiguousarray(rays, dtype=np.float32))
