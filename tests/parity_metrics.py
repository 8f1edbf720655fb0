"""Parity metrics of BASELINE.json's north_star, shared by the tests and the fixture generator.

check 1: primary-hit triangle/material IDs agree on >= 99.99 % of pixels, hit t within 1e-5 relative.
check 2: 1-spp radiance within 1e-3 relative per pixel on >= 99.9 % of pixels.
check 3: converged images: RMSE below 0.5 % of mean luminance.
"""
import numpy as np


def hits_agreement(t_a, tri_a, mat_a, t_b, tri_b, mat_b, rel=1e-5):
    """Fraction of pixels whose (triID.x, matID) agree, and, among those, whose t agrees to `rel`."""
    same_id = (tri_a == tri_b) & (mat_a == mat_b)
    denom = np.maximum(np.abs(t_b), 1e-30)
    close_t = np.abs(t_a - t_b) <= rel * denom
    return float(same_id.mean()), float((same_id & close_t).mean())


def radiance_agreement(a, b, rel=1e-3, abs_floor=1e-6):
    """Fraction of pixels where every channel of a is within `rel` (relative) of b.

    `abs_floor` keeps exactly-black reference pixels comparable: |a-b| <= rel*|b| + abs_floor.
    """
    a = np.asarray(a, np.float64).reshape(-1, 3)
    b = np.asarray(b, np.float64).reshape(-1, 3)
    ok = np.abs(a - b) <= rel * np.abs(b) + abs_floor
    return float(ok.all(axis=1).mean())


def luminance(img):
    img = np.asarray(img, np.float64).reshape(-1, 3)
    return 0.2126 * img[:, 0] + 0.7152 * img[:, 1] + 0.0722 * img[:, 2]


def rmse_over_mean_luminance(a, b):
    """RMSE of the luminance images divided by the mean luminance of b."""
    la, lb = luminance(a), luminance(b)
    return float(np.sqrt(np.mean((la - lb) ** 2)) / max(lb.mean(), 1e-30))
