"""Shared parity metrics: norm-relative error on matrices/vectors, max-relative on per-point steps."""
import numpy as np

TOL = 1e-4          # north_star: Hessian and update vector within 1e-4 relative (float)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    nb = np.linalg.norm(b)
    if nb == 0:
        return float(np.linalg.norm(a))
    return float(np.linalg.norm(a - b) / nb)


def max_rel(a, b, floor=1e-7):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
