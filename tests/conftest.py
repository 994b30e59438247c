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
                  if f.endswith(".npz") and f != "raygen.npz")


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
