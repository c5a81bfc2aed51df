"""Shared test inputs.  Nothing here reads /root/reference (absent on the GPU box)."""
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_json():
    return json.load(open(os.path.join(GOLDEN, "golden.json")))


def sha(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def ref_test1_inputs():
    """test/test1/test-setup.py:20-55 restated: A=C=1, B=1e-2 sin(2 pi r) sin(3 pi z), bc_init=0, 200x200 float32."""
    nr = nz = 200
    r = np.linspace(0.0, 1.0, nr); z = np.linspace(0.0, 1.0, nz)
    A = np.ones((nz, nr), np.float32); C = np.ones((nz, nr), np.float32)
    bc = np.zeros((nz, nr), np.float32)
    rr, zz = np.meshgrid(r, z)
    B = (1e-2 * np.sin(2.0 * np.pi * (rr - r[0]) / 1.0) * np.sin(3.0 * np.pi * (zz - z[0]) / 1.0)).astype(np.float32)
    return A, B, C, bc


def rel_l2(x, y):
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
    return float(np.sqrt(((x - y) ** 2).sum()) / max(np.sqrt((y ** 2).sum()), 1e-300))
