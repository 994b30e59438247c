"""CPU: the frame-assembly oracle (oracle/frame_oracle.py, scope row f-3) against the committed fixture produced
from the unmodified reference (oracle/make_golden_frame.py), and the embedded COLORMAP_JET table against OpenCV."""
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT
from oracle import frame_oracle as F

Z = np.load(os.path.join(GOLDEN_DIR, "frame_assembly.npz"))


@pytest.mark.parametrize("tag", ["blender", "llff", "raster"])
def test_frame_oracle_reproduces_golden(tag):
    H, W, s, near, far, with_gt = Z[f"{tag}_params"]
    gt = Z[f"{tag}_gt"] if with_gt else None
    with np.errstate(all="ignore"):
        u8, mat = F.assemble_frame(Z[f"{tag}_rgb"], Z[f"{tag}_depth"], int(H), int(W), int(s), float(near), float(far), gt)
    assert np.array_equal(u8, Z[f"{tag}_u8"]) and np.array_equal(mat, Z[f"{tag}_mat"])
    assert u8.shape == (int(H), int(W) * (3 if with_gt else 2), 3)


def test_numpy_cast_restatement():
    assert np.array_equal(F.astype_u8(Z["cast_in"]), Z["cast_out"])
    with np.errstate(all="ignore"):
        assert np.array_equal(Z["cast_in"].astype(np.uint8), Z["cast_out"])     # this host's numpy does the same


def test_unflatten_is_the_inverse_of_the_ray_grouping():
    H, W, s = 8, 12, 2
    raster = np.arange(H * W * 2, dtype=np.float32).reshape(H, W, 2)
    grouped = raster.reshape(H // s, s, W // s, s, 2).transpose(0, 2, 1, 3, 4).reshape(H * W, 2)   # '(h s1) (w s2) c -> (h w) (s1 s2) c'
    assert np.array_equal(F.unflatten_reshape(grouped, H, W, s), raster)


def test_embedded_jet_table_matches_opencv_and_fixture():
    src = open(os.path.join(ROOT, "nerf_sr_b200", "csrc", "nsr_jet_lut.h")).read()
    vals = [int(v, 16) for v in re.findall(r"0x([0-9a-f]{6})u", src)]
    assert len(vals) == 256
    table = np.array([[v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF] for v in vals], np.uint8)
    assert np.array_equal(table, Z["jet_lut"])
    assert np.array_equal(table, F.jet_lut())
