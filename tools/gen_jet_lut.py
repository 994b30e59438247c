"""Regenerate nerf_sr_b200/csrc/nsr_jet_lut.h from the installed OpenCV (cv2.COLORMAP_JET, the colormap the
reference's depth2im applies, utils/visualizer.py:164-176)."""
import os

import cv2
import numpy as np

lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_JET)[0]
rows = []
for i in range(0, 256, 8):
    rows.append("  " + ", ".join(f"0x{(int(lut[j, 0]) | (int(lut[j, 1]) << 8) | (int(lut[j, 2]) << 16)):06x}u" for j in range(i, i + 8)) + ",")
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nerf_sr_b200", "csrc", "nsr_jet_lut.h")
src = open(path).read()
head = src[:src.index("static const uint32_t kJetLut[256] = {")]
open(path, "w").write(head + "static const uint32_t kJetLut[256] = {\n" + "\n".join(rows) + "\n};\n")
print("wrote", path)
