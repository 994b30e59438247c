"""Import the UNMODIFIED reference (cwchenwang/NeRF-SR) in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/nerf_oracle.py header).  /root/reference
does not exist on the GPU box; ``smoke()`` and ``bench.py`` never import this
module.  It is used by ``oracle/make_golden*.py``, by the CPU-only test that
re-pins the oracle when the reference tree is present, and by
``tests/test_gpu_reference_model.py``, which runs the reference's own model
class under ``patch_model`` on the B200 from the git-ignored copy that
``tools/stage_reference.py`` packs into ``baseline/_ref/NeRF-SR.tar.gz`` (the
tests skip when no archive travelled).

The reference does not import as-is here: numpy 2.x
dropped ``numpy.lib.shape_base``, ``dominate`` and ``imageio`` are absent, and
``BaseOptions.parse`` calls ``torch.cuda.set_device``.  We pre-seed
``sys.modules`` with inert stubs and build ``opt`` by hand from the model's own
``modify_commandline_options`` -- the reference tree is never modified.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

_ARCHIVE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "NeRF-SR.tar.gz")


def _resolve_root() -> str:
    env = os.environ.get("NSR_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/models/nerf_downX_model.py"):
        return "/root/reference"
    if os.path.isfile(_ARCHIVE):                      # the GPU box: unpack the staged archive once per process tree
        import hashlib
        import tarfile
        import tempfile
        tag = hashlib.sha1(open(_ARCHIVE, "rb").read()).hexdigest()[:12]
        dst = os.path.join(tempfile.gettempdir(), f"nsr_reference_{tag}")
        if not os.path.isfile(os.path.join(dst, "models", "nerf_downX_model.py")):
            tmp = tempfile.mkdtemp(prefix="nsr_reference_")
            with tarfile.open(_ARCHIVE) as tar:
                tar.extractall(tmp, filter="data")
            try:
                os.rename(tmp, dst)
            except OSError:                           # another process won the race
                pass
        return dst
    return "/root/reference"


REFERENCE_ROOT = _resolve_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "nerf_downX_model.py"))


def _install_stubs() -> None:
    import numpy as np
    if "numpy.lib.shape_base" not in sys.modules:
        m = types.ModuleType("numpy.lib.shape_base")
        m.expand_dims = np.expand_dims
        sys.modules["numpy.lib.shape_base"] = m
    if "dominate" not in sys.modules:
        dom = types.ModuleType("dominate")
        tags = types.ModuleType("dominate.tags")
        for n in ("meta", "h3", "table", "tr", "td", "p", "a", "img", "br"):
            setattr(tags, n, lambda *a, **k: None)
        dom.tags = tags
        dom.document = lambda *a, **k: None
        sys.modules["dominate"] = dom
        sys.modules["dominate.tags"] = tags
    if "imageio" not in sys.modules:
        sys.modules["imageio"] = types.ModuleType("imageio")


def load_reference_model(model_name: str = "nerf_downX", extra_args=(), device: str = "cpu", train: bool = False):
    """Return (model, opt): a reference NeRFDownXModel / NeRFModel on ``device``
    in eval mode with kaiming-initialised nets (caller overwrites weights).
    train=True builds it the way train.py does (isTrain: the model creates its Adam optimiser,
    models/nerf_downX_model.py:198-204) and leaves it in train mode."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torch
    if model_name == "nerf_downX":
        from models.nerf_downX_model import NeRFDownXModel as Model
    elif model_name == "nerf":
        from models.nerf_model import NeRFModel as Model
    else:
        raise ValueError(model_name)

    parser = argparse.ArgumentParser()
    # the handful of base flags the model reads (options/base_options.py:35-74)
    parser.add_argument("--accelerator", default="dp")
    parser.add_argument("--name", default="oracle")
    parser.add_argument("--checkpoints_dir", default="/tmp/nsr_oracle_ckpt")
    parser.add_argument("--init_type", default="kaiming")
    parser.add_argument("--init_gain", type=float, default=0.02)
    parser.add_argument("--sisr_path", default=None)
    parser.add_argument("--img_wh", type=int, nargs=2, default=[8, 8])
    parser.add_argument("--patch_size", type=int, default=1)
    parser.add_argument("--ray_chunk", type=int, default=4096)
    parser.add_argument("--point_chunk", type=int, default=2048 * 128)
    parser.add_argument("--lr", type=float, default=5e-4)
    parser.add_argument("--beta1", type=float, default=0.9)
    # train flags read by optimize_parameters / the schedulers (options/train_options.py:33-55)
    parser.add_argument("--grad_clip_val", type=float, default=0)
    parser.add_argument("--grad_clip_type", type=str, default="norm")
    parser.add_argument("--lr_policy", type=str, default="exp")
    parser.add_argument("--lr_final", type=float, default=5e-6)
    parser.add_argument("--n_epochs", type=int, default=20)
    parser.add_argument("--n_epochs_decay", type=int, default=10)
    parser = Model.modify_commandline_options(parser)
    opt = parser.parse_args(list(extra_args))
    opt.isTrain, opt.isTest, opt.isInfer = (True, False, False) if train else (False, True, False)
    opt.device = torch.device(device)
    opt.n_gpus = 0 if device == "cpu" else 1
    opt.gpu_ids = [] if device == "cpu" else [0]
    opt.is_master = True
    model = Model(opt)
    if train:
        model.train()
    else:
        model.eval()
    return model, opt


def set_weights(model, p_coarse, p_fine) -> None:
    """Load oracle-style parameter dicts into the reference nets."""
    import torch
    for net, p in ((model.netCoarse, p_coarse), (model.netFine, p_fine)):
        target = net.module if hasattr(net, "module") else net
        target.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
