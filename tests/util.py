"""Shared helpers for the parity tests."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


def same_bits(a, b):
    """bit-for-bit equality of two float64 arrays (NaNs must coincide, payloads ignored)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return np.array_equal(bits(np.where(na, 0.0, a)), bits(np.where(nb, 0.0, b)))


def n_mismatch(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    return int(np.count_nonzero((na != nb) | (bits(np.where(na, 0.0, a)) != bits(np.where(nb, 0.0, b)))))


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


OPERATOR_CASES = ["2d_circle_32", "2d_sflow_40x24", "2d_zalesak_64", "2d_leveque_32",
                  "3d_sphere_16", "3d_sflow_20x12x16", "3d_leveque_16", "3d_leveque_32"]
