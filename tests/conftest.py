"""Shared test helpers.  GPU tests are marked ``@pytest.mark.gpu``; everything
else must pass on a CPU-only box (``pytest -m "not gpu"``)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith(".npz") and f not in ("raygen.npz", "frame_assembly.npz", "pose_paths.npz")
                  and not f.startswith("train_step_") and not f.startswith("scene_"))


def materialize_scene(name, dest):
    """Write the raw capture files stored in tests/golden/<name>.npz (COLMAP binaries / transforms JSON / PNGs, the
    inputs oracle/make_golden_scenes.py fed to the reference's dataset classes) under ``dest``; returns (npz, meta)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    for key in z.files:
        if key.startswith("file_"):
            path = os.path.join(dest, key[len("file_"):])
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "wb") as fh:
                fh.write(z[key].tobytes())
    return z, json.loads(bytes(z["meta_json"]).decode())


def train_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("train_step_") and f.endswith(".npz"))


class TrainFixture:
    """A committed training-step fixture (tests/golden/train_step_*.npz, produced by
    oracle/make_golden_train.py from the unmodified reference's optimize_parameters)."""

    def __init__(self, name):
        from oracle import nerf_oracle as O
        from oracle import train_oracle as T
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name, self.z = name, z
        self.meta = json.loads(bytes(z["meta_json"]).decode())
        self.cfg = O.RenderConfig(**self.meta["cfg"])
        self.tcfg = T.TrainConfig(**self.meta["tcfg"])
        self.s = int(self.meta["s"])
        self.rays = torch.from_numpy(z["rays"])
        self.target = torch.from_numpy(z["target"])
        self.target_sr = torch.from_numpy(z["target_sr"]) if "target_sr" in z.files else None   # data_rgbs_sr (--sisr_path)
        # --with_ref: the reference-view batch (data_ref_rays / data_ref_rgbs) and its own train-mode draws
        self.ref_rays = torch.from_numpy(z["ref_rays"]) if "ref_rays" in z.files else None
        self.ref_rgbs = torch.from_numpy(z["ref_rgbs"]) if "ref_rgbs" in z.files else None
        self.rng, self.ref_rng = [], []
        for step in range(2):
            get = lambda f: torch.from_numpy(z[f"rng{step}_{f}"]) if f"rng{step}_{f}" in z.files else None
            self.rng.append(O.RenderRng(get("u_coarse"), get("noise_coarse"), get("u_fine"), get("noise_fine")))
            getr = lambda f: torch.from_numpy(z[f"refrng{step}_{f}"]) if f"refrng{step}_{f}" in z.files else None
            self.ref_rng.append(None if self.ref_rays is None else
                                O.RenderRng(getr("u_coarse"), getr("noise_coarse"), getr("u_fine"), getr("noise_fine")))
        sd = self.meta["seeds"]
        self.p_coarse = O.make_mlp_params(self.cfg, sd[0])
        self.p_fine = O.make_mlp_params(self.cfg, sd[1])

    def sub(self, prefix, step, i):
        return torch.from_numpy(self.z[f"{prefix}{step}_{i}_sub"]), float(self.z[f"{prefix}{step}_{i}_norm"][0])


def rng_dict(rng):
    if rng is None:
        return None
    return {k: getattr(rng, k) for k in ("u_coarse", "noise_coarse", "u_fine", "noise_fine") if getattr(rng, k) is not None}


class Fixture:
    """A committed golden fixture (tests/golden/<name>.npz, produced by
    oracle/make_golden.py from the unmodified reference)."""

    def __init__(self, name):
        from oracle import nerf_oracle as O
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.meta = json.loads(bytes(z["meta_json"]).decode())
        cfgd = dict(self.meta["cfg"])
        if "skips" in cfgd:
            cfgd["skips"] = tuple(cfgd["skips"])
        self.cfg = O.RenderConfig(**cfgd)
        self.rays = torch.from_numpy(z["rays"])
        self.out = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("out_")}
        self.z_coarse = torch.from_numpy(z["z_coarse"])
        self.raw_coarse = torch.from_numpy(z["raw_coarse"])
        self.z_fine = torch.from_numpy(z["z_fine"]) if "z_fine" in z.files else None
        self.raw_fine = torch.from_numpy(z["raw_fine"]) if "raw_fine" in z.files else None
        self.rng = None
        if self.meta["train"]:
            get = lambda f: torch.from_numpy(z["rng_" + f]) if ("rng_" + f) in z.files else None
            self.rng = O.RenderRng(get("u_coarse"), get("noise_coarse"), get("u_fine"), get("noise_fine"))
        s = self.meta["seeds"]
        self.p_coarse = O.make_mlp_params(self.cfg, s[0], self.meta["sigma_bias"], self.meta["bias_std"])
        self.p_fine = O.make_mlp_params(self.cfg, s[1], self.meta["sigma_bias"], self.meta["bias_std"])


@pytest.fixture(scope="session")
def load_fixture():
    cache = {}

    def _load(name):
        if name not in cache:
            cache[name] = Fixture(name)
        return cache[name]
    return _load


# ---- end-to-end fine-stage acceptance rule (SURVEY.md 8c iii), shared by the GPU parity tests -------------------------------
# The fine pass places its samples by an inverse-CDF bin search over the coarse weights: a decision that flips under
# perturbations of 1e-6, after which 2^9-frequency encodings amplify the shift.  Two correct evaluations therefore
# disagree on a small FRACTION of fine outputs; what can be tested is that our fraction is that of an fp32-grade evaluation.
# floor = violation fraction of the reference's own fp32 output against the same computation in fp64 on the same rays.
#   ours vs fp64   <= 2 floor + margin(precision) + 3 sigma(n_rays)
#   ours vs ref32  <= 2 floor + margin(precision) + 3 sigma(n_rays)
# (2 floor: flipped decisions add up -- ours-vs-fp64 can be as large as (ours vs ref32) + (ref32 vs fp64), and two independent
#  fp32-grade evaluations each sit one floor away from the exact answer; measured on full frames, profiles/r02_frame_parity.md:
#  fp16x3 1.3-1.4 x floor against ref32, bf16x3 1.7 x floor on Blender-like and 5.5 x a 0.2 % floor on LLFF-like rays.)
# margin: percentage points we allow above the reference's own arithmetic -- 0.01 for the paths whose operands carry fp32-grade
# mantissas (fp32 CUDA cores; fp16 hi/lo = 22 bits), 0.02 for bf16 hi/lo (16 bits: ~1e-5 relative error on the coarse weights
# instead of ~1e-6, which moves proportionally more bin decisions; measured 1.0-1.4 % vs a floor of 0.1-0.25 % on LLFF-like rays,
# 2.4 % vs 1.1 % on Blender-like rays, profiles/r02_frame_parity.md).  sigma(n) = sqrt(p (1 - p) / n_rays), p = max(floor, 0.01):
# the binomial sampling error of a violation COUNT on a fixture of n_rays rays (0.02 on the 128-ray golden fixtures, < 0.001 on
# full frames) -- it replaces round 1's flat 0.06 cushion and vanishes where the statistics allow a tight statement.
E2E_MARGIN = {"bf16x3": 0.02, "fp16x3": 0.01, "fp32_simt": 0.01, "bf16": 1.0}


def e2e_bounds(floor: float, n_rays: int, prec: str):
    p = max(float(floor), 0.01)
    sigma3 = 3.0 * (p * (1.0 - p) / max(int(n_rays), 1)) ** 0.5
    return 2.0 * floor + E2E_MARGIN[prec] + sigma3, 2.0 * floor + E2E_MARGIN[prec] + sigma3
