"""Geometric verification on libdimb200 - the estimator of the reference's ``utils/geometric_verification.py:45-179`` on the GPU.

``geometric_verification(kpts0, kpts1, method, threshold, confidence, max_iters)`` keeps the reference's signature and return
contract ``(F, inlMask)``: ``F`` a (3,3) fundamental matrix or ``None``, ``inlMask`` a boolean array over the correspondences;
method NONE returns ``(None, all True)`` (:100-101) and fewer than 8 matches return ``(None, all True)`` (:107-111).  Every other
method name of the reference (PYDEGENSAC, MAGSAC, RANSAC, the OpenCV USAC family) selects ONE estimator here: seeded 8-point
RANSAC with Sampson inliers and two least-squares refits (csrc/gv.cu), all hypotheses evaluated in parallel.  ``confidence`` is
accepted and unused (there is no adaptive stopping: min(max_iters, 8192) hypotheses always run).  Like the reference's
estimators the result is stochastic in the sense that it depends on the seed; parity is statistical (tests/test_geometry.py).
"""
from __future__ import annotations

import numpy as np

METHODS = ("NONE", "PYDEGENSAC", "MAGSAC", "RANSAC", "LMEDS", "RHO", "USAC_DEFAULT", "USAC_PARALLEL", "USAC_FM_8PTS", "USAC_FAST",
           "USAC_ACCURATE", "USAC_PROSAC", "USAC_MAGSAC")


def geometric_verification(kpts0: np.ndarray = None, kpts1: np.ndarray = None, method="pydegensac", threshold: float = 1, confidence: float = 0.9999,
                           max_iters: int = 10000, quiet: bool = False, device: int = 0, seed: int = 0, **kwargs):
    from . import _native
    name = getattr(method, "name", method)
    if isinstance(name, int):
        name = METHODS[name] if 0 <= name < len(METHODS) else None
    if not isinstance(name, str) or name.upper() not in METHODS:
        raise ValueError(f"Invalid Geometry Verification method. It must be one of {list(METHODS)}")
    n = len(kpts0)
    if name.upper() == "NONE" or n < 8:
        return None, np.ones(n, dtype=bool)
    return _native.Context.get(device).gv_fundamental(kpts0, kpts1, threshold, max_iters, seed)
