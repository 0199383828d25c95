"""Loader of the frozen golden inputs (tests/golden/stream_inputs.npz, written by make_golden.py)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4"),
                        ("b", "u1"), ("g", "u1"), ("r", "u1"), ("a", "u1"), ("pad", "u1", (12,))])
_npz = None


def frozen_names():
    global _npz
    if _npz is None:
        _npz = np.load(os.path.join(HERE, "stream_inputs.npz"))
    return sorted(k[:-4] for k in _npz.files if k.endswith(".xyz"))


def load_case(name):
    """The 32-byte PointXYZRGB records of a frozen case: x,y,z, 1.0f, b,g,r, a=255, zero padding."""
    frozen_names()
    xyz, bgr = _npz[name + ".xyz"], _npz[name + ".bgr"]
    p = np.zeros(xyz.shape[0], POINT_DTYPE)
    p["x"], p["y"], p["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    p["w"] = 1.0
    p["b"], p["g"], p["r"] = bgr[:, 0], bgr[:, 1], bgr[:, 2]
    p["a"] = 255
    return p
