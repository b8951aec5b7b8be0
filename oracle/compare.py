"""Parity comparison helpers (test infrastructure only).

Index outputs must be bit-exact *except* where the reference's own decision is within arithmetic noise of a
threshold: the top-k cut of SuperPoint (K-th vs (K+1)-th score) and the LightGlue filter threshold.  Such
differences are never dropped silently: they are accepted only if every differing item sits within `tol` of the
decision boundary in the oracle's own scores, and they are returned to the caller for reporting.
"""
from __future__ import annotations

import numpy as np


def compare_superpoint(out: dict, ref: dict, ref_nms: np.ndarray | None = None, tol: float = 1e-4, max_boundary: int = 8) -> dict:
    """out/ref: FeaturesDicts. ref_nms: oracle's NMS score map (needed only if the keypoint sets differ)."""
    ko = [tuple(k) for k in out["keypoints"].astype(np.int64)]
    kr = [tuple(k) for k in ref["keypoints"].astype(np.int64)]
    so, sr = set(ko), set(kr)
    assert len(so) == len(ko) and len(sr) == len(kr), "duplicate keypoints"
    assert len(ko) == len(kr), f"keypoint count {len(ko)} != {len(kr)}"
    diff = so ^ sr
    if diff:
        assert ref_nms is not None, f"keypoint sets differ in {len(diff)} points"
        assert len(diff) <= max_boundary, f"{len(diff)} keypoints differ"
        cut = float(ref["scores"].min())  # K-th largest score of the oracle
        for (x, y) in diff:
            s = float(ref_nms[y, x])
            assert s > 0 and abs(s - cut) < tol, f"keypoint {(x, y)} differs and is not at the top-k cut: score {s} vs cut {cut}"
    io = {k: i for i, k in enumerate(ko)}
    ir = {k: i for i, k in enumerate(kr)}
    common = sorted(so & sr)
    a = np.array([io[k] for k in common], np.int64)
    b = np.array([ir[k] for k in common], np.int64)
    ds = float(np.abs(out["scores"][a] - ref["scores"][b]).max()) if len(a) else 0.0
    dd = float(np.abs(out["descriptors"][:, a] - ref["descriptors"][:, b]).max()) if len(a) else 0.0
    assert ds < tol, f"score error {ds}"
    assert dd < tol, f"descriptor error {dd}"
    return {"n": len(ko), "boundary_diffs": sorted(diff), "max_dscore": ds, "max_ddesc": dd}


def compare_matches(out: dict, ref: dict, filter_threshold: float = 0.1, tol: float = 1e-4, max_boundary: int = 4) -> dict:
    """out/ref: dict(matches (S,2), scores (S,), stop). Differences only allowed at the filter threshold."""
    assert out["stop"] == ref["stop"], f"stop layer {out['stop']} != {ref['stop']}"
    mo = {tuple(m): float(s) for m, s in zip(out["matches"], out["scores"])}
    mr = {tuple(m): float(s) for m, s in zip(ref["matches"], ref["scores"])}
    diff = set(mo) ^ set(mr)
    assert len(diff) <= max_boundary, f"{len(diff)} matches differ"
    for m in diff:
        s = mo.get(m, mr.get(m))
        assert abs(s - filter_threshold) < tol, f"match {m} differs and is not at the filter threshold (score {s})"
    common = sorted(set(mo) & set(mr))
    ds = max((abs(mo[m] - mr[m]) for m in common), default=0.0)
    assert ds < tol, f"match score error {ds}"
    if len(out["matches"]) > 1:
        assert np.all(np.diff(out["matches"][:, 0]) > 0), "matches not ascending in index 0"
    return {"n": len(mr), "boundary_diffs": sorted(diff), "max_dscore": ds}
