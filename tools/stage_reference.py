"""Stage the UNMODIFIED reference tree for the GPU box (test infrastructure only).

/root/reference exists only in the build container.  The GPU tests that run the reference's own
`NeRFDownXModel` class under `patch_model` (tests/test_gpu_reference_model.py) need its Python files on the
B200 box, so `__graft_entry__.build()` packs them -- minus docs/ and .git -- into ONE archive,
`baseline/_ref/NeRF-SR.tar.gz`: git-ignored (never part of this repo's history), not gpurun-ignored (it
travels with the snapshot), in the directory the bench contract reserves for the installed reference.
`oracle/ref_shim.py` unpacks it into a temporary directory when /root/reference is absent.  Nothing in the
product path reads it."""
from __future__ import annotations

import os
import sys
import tarfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("NSR_REFERENCE_SRC", "/root/reference")
ARCHIVE = os.path.join(ROOT, "baseline", "_ref", "NeRF-SR.tar.gz")
KEEP_DIRS = ("models", "utils", "data", "options", "scripts")


def stage(force: bool = False) -> str | None:
    if not os.path.isfile(os.path.join(SRC, "models", "nerf_downX_model.py")):
        return ARCHIVE if os.path.isfile(ARCHIVE) else None
    if os.path.isfile(ARCHIVE) and not force:
        return ARCHIVE
    os.makedirs(os.path.dirname(ARCHIVE), exist_ok=True)

    def keep(ti: tarfile.TarInfo):
        return None if ("__pycache__" in ti.name or ti.name.endswith(".pyc")) else ti
    with tarfile.open(ARCHIVE, "w:gz") as tar:
        for d in KEEP_DIRS:
            if os.path.isdir(os.path.join(SRC, d)):
                tar.add(os.path.join(SRC, d), arcname=d, filter=keep)
        for f in sorted(os.listdir(SRC)):
            if f.endswith(".py") or f in ("requirements.txt", "README.md"):
                tar.add(os.path.join(SRC, f), arcname=f)
    return ARCHIVE


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
