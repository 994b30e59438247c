"""CPU-only host-logic tests: the C-ABI library loads and exports every declared symbol, the
binding's structs match the header, and it refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT
from oracle import nerf_oracle as O


def _build():
    import __graft_entry__ as g
    g.build()


def test_library_builds_and_exports_every_declared_symbol():
    _build()
    from nerf_sr_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "nsr.h")).read()
    declared = set(re.findall(r"NSR_API [a-z0-9_ \*]+?\b(nsr_[a-z_]+)\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.nsr_abi_version() == _lib.ABI_VERSION


def test_config_struct_matches_header():
    from nerf_sr_b200 import _lib
    header = open(os.path.join(ROOT, "include", "nsr.h")).read()
    body = header[header.index("typedef struct NsrConfig {"):header.index("} NsrConfig;")]
    fields = re.findall(r"^\s*(?:uint32_t|int32_t|float)\s+([a-z_A-Z]+)(?:\[\d+\])?;", body, flags=re.M)
    assert fields == [f[0] for f in _lib.NsrConfig._fields_]
    assert C.sizeof(_lib.NsrConfig) == 4 * (len(fields) - 1) + 4 * 8


def test_header_is_plain_c_and_struct_sizes_match_the_binding(tmp_path):
    """include/nsr.h must compile as C99 on its own (it is what a cgo / JNI / ctypes maintainer binds), and every
    struct the ctypes stub mirrors must have the same size and field offsets as the C compiler's."""
    import shutil
    import subprocess
    from nerf_sr_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    structs = {"NsrConfig": _lib.NsrConfig, "NsrRng": _lib.NsrRng, "NsrOutputs": _lib.NsrOutputs, "NsrPassOutputs": _lib.NsrPassOutputs,
               "NsrOutGrads": _lib.NsrOutGrads, "NsrLossTerms": _lib.NsrLossTerms, "NsrRayGen": _lib.NsrRayGen, "NsrLrOutputs": _lib.NsrLrOutputs}
    lines = ['#include "nsr.h"', "#include <stdio.h>", "#include <stddef.h>", "int main(void) {"]
    for name, cls in structs.items():
        lines.append(f'  printf("{name} %zu", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({name}, {field}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src = os.path.join(str(tmp_path), "abi.c")
    open(src, "w").write("\n".join(lines))
    exe = os.path.join(str(tmp_path), "abi")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    assert len(out) == len(structs)
    for line in out:
        name, size, *offsets = line.split()
        cls = structs[name]
        assert int(size) == C.sizeof(cls), name
        assert [int(o) for o in offsets] == [getattr(cls, f).offset for f, _ in cls._fields_], name


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    _build()
    from nerf_sr_b200 import NsrError, Renderer, _lib, config_from_opt
    with pytest.raises(NsrError):
        Renderer(O.RenderConfig())
    lib = _lib.load()
    cfg = config_from_opt(O.RenderConfig(), 0)
    h = C.c_void_p()
    assert lib.nsr_create(C.byref(cfg), C.byref(h)) == 6          # NSR_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.nsr_last_error(None)
    cfg.struct_size = 12                                             # ABI skew is caught first
    assert lib.nsr_create(C.byref(cfg), C.byref(h)) == 1


def test_config_from_reference_style_opt():
    from types import SimpleNamespace
    from nerf_sr_b200 import config_from_opt
    opt = SimpleNamespace(D=8, W=256, skips=[4], N_coarse=64, N_importance=64, white_bkgd=True, noise_std=1.0,
                          sigma_activation="softplus", color_activation="none", lindisp=True, deg_pos=10, deg_dir=4)
    c = config_from_opt(opt, 0, "fp16x3", viewdir_offset=8)
    assert (c.skips_mask, c.white_bkgd, c.sigma_activation, c.color_activation, c.lindisp) == (16, 1, 1, 1, 1)
    assert c.precision == 2 and c.viewdir_offset == 8 and abs(c.noise_std - 1.0) < 1e-7


def test_state_dict_order_matches_oracle_shapes():
    from nerf_sr_b200 import state_dict_order
    assert state_dict_order(8) == [n for n, _ in O.mlp_param_shapes(O.RenderConfig())]


def test_pose_paths_match_reference_golden():
    """Scope row f-4: nerf_sr_b200/paths.py against outputs of the reference's own generators
    (oracle/make_golden_paths.py; data/llff_downX_dataset.py:20-160)."""
    import numpy as np
    from conftest import GOLDEN_DIR
    from nerf_sr_b200 import paths as P
    z = np.load(os.path.join(GOLDEN_DIR, "pose_paths.npz"))
    rx, ry, rz, focus, n = z["spiral_in"]
    assert np.allclose(P.spiral_poses(np.array([rx, ry, rz]), float(focus), int(n)), z["spiral"], rtol=0, atol=1e-14)
    radius, n = z["spheric_in"]
    sp = P.spheric_poses(float(radius), int(n))
    assert np.allclose(sp, z["spheric"], rtol=0, atol=1e-14)
    assert sp.shape == (int(n), 3, 4) and np.allclose(np.linalg.det(sp[:, :, :3]), 1.0)    # proper rotations
    centered, avg = P.center_poses(z["poses"])
    assert np.allclose(centered, z["centered"], rtol=0, atol=1e-13) and np.allclose(avg, z["avg"], rtol=0, atol=1e-14)
    c2, a2 = P.center_poses(centered)              # centring is idempotent: the average of centred poses is the identity frame
    assert np.allclose(a2[:, :3], np.eye(3), atol=1e-12) and np.allclose(a2[:, 3], 0, atol=1e-12)


def test_oracle_is_imported_only_by_the_checkers():
    """oracle/ is test infrastructure: nothing under nerf_sr_b200/ or tools/ may import it, and bench.py only inside its
    baseline arms (cpu_baseline / --impl reference / the torch-on-GPU reference port)."""
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    for sub in ("nerf_sr_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not pat.search(src), os.path.join(dirpath, f)
    bench = open(os.path.join(ROOT, "bench.py")).read()
    top_level = [m for m in pat.finditer(bench) if not m.group(0).startswith((" ", "\t"))]
    assert not top_level, "bench.py imports the oracle at module level"
    import ast
    tree = ast.parse(bench)
    allowed = {"cpu_reference_rate", "torch_gpu_reference_rate", "oracle_train_rate", "run_reference", "run_reference_train"}
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = any(isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] == "oracle" for n in ast.walk(fn))
        if uses:
            assert fn.name in allowed, f"bench.py:{fn.name} imports the oracle outside a baseline arm"


def test_missing_library_fails_loudly(tmp_path):
    """No silent fallback when libnsr_b200.so is absent: importing the binding with a bad path raises, naming the build step."""
    import subprocess
    import sys
    code = ("import nerf_sr_b200._lib as L\n"
            "try:\n    L.load()\nexcept ImportError as e:\n    print('IMPORT_ERROR', e)\nelse:\n    print('LOADED')\n")
    env = dict(os.environ, NSR_LIB_PATH=os.path.join(str(tmp_path), "nowhere", "libnsr_b200.so"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-1500:]
    assert "IMPORT_ERROR" in p.stdout and "no CPU or PyTorch fallback" in p.stdout and "build" in p.stdout
