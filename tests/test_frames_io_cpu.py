"""CPU: the test-sweep output files (nerf_sr_b200/frames.py, scope row f-3): PNG bytes decode to the pixels the
reference's ``_save_image`` stores, depth matrices load the way warp.py reads them, file names follow ``test()``."""
import os

import numpy as np
import pytest
import torch

from nerf_sr_b200 import frames as F
from oracle import ref_shim


def _decode_cv2(path):
    import cv2
    return cv2.cvtColor(cv2.imread(path, cv2.IMREAD_UNCHANGED), cv2.COLOR_BGR2RGB)


@pytest.mark.parametrize("shape", [(7, 13, 3), (1, 1, 3), (64, 200, 3), (5, 9)])
def test_png_roundtrip(tmp_path, shape):
    from PIL import Image
    g = np.random.default_rng(sum(shape))
    img = g.integers(0, 256, shape, dtype=np.uint8)
    p = os.path.join(str(tmp_path), "x.png")
    F.write_png(p, img)
    assert np.array_equal(np.asarray(Image.open(p)), img)
    if img.ndim == 3:
        assert np.array_equal(_decode_cv2(p), img)
    # smooth images compress (the row filter is doing its job) and stay exact
    ramp = np.broadcast_to(np.arange(200, dtype=np.uint8)[None, :, None], (64, 200, 3)).copy()
    data = F.encode_png(ramp)
    assert len(data) < ramp.size // 20
    open(p, "wb").write(data)
    assert np.array_equal(np.asarray(Image.open(p)), ramp)


def test_png_rejects_bad_input():
    with pytest.raises(ValueError):
        F.encode_png(np.zeros((4, 4, 3), dtype=np.float32))
    with pytest.raises(ValueError):
        F.encode_png(np.zeros((0, 4, 3), dtype=np.uint8))
    with pytest.raises(ValueError):
        F.encode_png(np.zeros((4, 4, 2), dtype=np.uint8))


def test_depth_npz_is_what_warp_reads(tmp_path):
    mat = np.random.default_rng(1).random((6, 8)).astype(np.float32)
    mat[2, 3] = np.nan
    p = os.path.join(str(tmp_path), "0-fine-depth-ori.npz")
    F.write_depth_npz(p, torch.from_numpy(mat))
    got = np.load(p)["arr_0"]                      # warp.py:112
    assert got.dtype == np.float32 and got[2, 3] == 0 and np.array_equal(got, np.nan_to_num(mat))


def test_save_test_sweep_names_and_contents(tmp_path):
    g = np.random.default_rng(2)

    def frame():
        return {"fine_pred": torch.from_numpy(g.integers(0, 256, (4, 12, 3), dtype=np.uint8)),
                "fine_pred_ori": torch.from_numpy(g.integers(0, 256, (8, 24, 3), dtype=np.uint8)),
                "fine_depth_mat_ori": torch.from_numpy(g.random((8, 12)).astype(np.float32))}
    frames = [frame() for _ in range(3)]
    n = F.save_test_sweep(str(tmp_path), iter(frames), gif_keys=("fine_pred",))
    assert n == 3
    assert sorted(os.listdir(str(tmp_path))) == sorted(
        [f"{i}-fine.png" for i in range(3)] + [f"{i}-fine-ori.png" for i in range(3)]
        + [f"{i}-fine-depth-ori.npz" for i in range(3)] + ["fine.gif"])
    assert np.array_equal(_decode_cv2(os.path.join(str(tmp_path), "1-fine-ori.png")), frames[1]["fine_pred_ori"].numpy())
    assert np.array_equal(np.load(os.path.join(str(tmp_path), "2-fine-depth-ori.npz"))["arr_0"], frames[2]["fine_depth_mat_ori"].numpy())
    from PIL import Image
    gif = Image.open(os.path.join(str(tmp_path), "fine.gif"))
    assert gif.n_frames == 3 and gif.size == (12, 4)
    with pytest.raises(KeyError):
        F.save_test_frame(str(tmp_path), 0, frames[0], keys=["fine_weights"])


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree only exists in the build container")
def test_png_and_npz_equal_the_reference_savers(tmp_path):
    """utils/visualizer.py's own _save_image / _save_matrix on the same panel: identical decoded pixels / matrix."""
    import sys
    ref_shim._install_stubs()
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    from utils import visualizer as V
    g = np.random.default_rng(3)
    panel = g.random((10, 30, 3)).astype(np.float32)
    ref_dir, my_dir = os.path.join(str(tmp_path), "ref"), os.path.join(str(tmp_path), "mine")
    os.makedirs(ref_dir), os.makedirs(my_dir)
    mat = g.random((10, 15)).astype(np.float32)
    V.save_visuals(ref_dir, {"a": V.Visualizee("image", panel, timestamp=False, name="3-fine-ori", data_format="HWC", range=(0, 1),
                                               img_format="png"),
                             "b": [V.Visualizee("matrix", mat, timestamp=False, name="3-fine-depth-ori")]})
    F.save_test_frame(my_dir, 3, {"fine_pred_ori": (panel * 255.0).astype(np.uint8), "fine_depth_mat_ori": mat})
    assert np.array_equal(_decode_cv2(os.path.join(ref_dir, "3-fine-ori.png")), _decode_cv2(os.path.join(my_dir, "3-fine-ori.png")))
    assert np.array_equal(np.load(os.path.join(ref_dir, "3-fine-depth-ori.npz"))["arr_0"],
                          np.load(os.path.join(my_dir, "3-fine-depth-ori.npz"))["arr_0"])


def test_png_roundtrip_property():
    """Arbitrary shapes and contents (hypothesis): what PIL decodes is what was encoded."""
    import io
    from hypothesis import given, settings, strategies as st
    from PIL import Image

    @settings(max_examples=40, deadline=None)
    @given(st.integers(1, 40), st.integers(1, 40), st.booleans(), st.integers(0, 2 ** 31 - 1), st.sampled_from(["noise", "flat", "ramp"]))
    def check(h, w, colour, seed, kind):
        g = np.random.default_rng(seed)
        shape = (h, w, 3) if colour else (h, w)
        if kind == "noise":
            img = g.integers(0, 256, shape, dtype=np.uint8)
        elif kind == "flat":
            img = np.full(shape, g.integers(0, 256), dtype=np.uint8)
        else:
            img = (np.arange(int(np.prod(shape))) % 256).astype(np.uint8).reshape(shape)
        got = np.asarray(Image.open(io.BytesIO(F.encode_png(img))))
        assert got.shape == img.shape and np.array_equal(got, img)
    check()
