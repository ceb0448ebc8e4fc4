"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def rel_err(a, b):
    """||a-b||_2 / ||b||_2 over the whole array, float64 (the 'relative' of the stated tolerances)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.linalg.norm(b.ravel())
    num = np.linalg.norm((a - b).ravel())
    return num / den if den > 0 else num


# unit roundoff of the storage formats (round-to-nearest): half an ulp relative
EPS = {"float32": 2.0 ** -24, "fp32": 2.0 ** -24, "bfloat16": 2.0 ** -8, "bf16": 2.0 ** -8,
       "float16": 2.0 ** -11, "fp16": 2.0 ** -11}
