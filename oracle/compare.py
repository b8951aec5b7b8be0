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


def compare_aliked(out: dict, ref: dict, ref_score_map: np.ndarray | None = None, threshold: float = 0.2, radius: int = 2,
                   tol: float = 1e-4, tol_kpt: float = 1e-3, max_boundary: int = 4) -> dict:
    """ALIKED FeaturesDicts (sub-pixel keypoints): pair keypoints by nearest neighbour.  Unpaired keypoints are
    tolerated only where the oracle's score map holds, within the NMS radius, a value within ``tol`` of the detection
    threshold or of the n_limit cut (the decision is then within fp32 summation-order noise)."""
    ko, kr = out["keypoints"].astype(np.float64), ref["keypoints"].astype(np.float64)
    assert out["descriptors"].shape[0] == ref["descriptors"].shape[0]
    if len(kr) == 0 or len(ko) == 0:
        assert len(ko) == len(kr), f"keypoint count {len(ko)} != {len(kr)}"
        return {"n": 0, "boundary_diffs": [], "max_dkpt": 0.0, "max_dscore": 0.0, "max_ddesc": 0.0}
    d = np.linalg.norm(ko[:, None] - kr[None], axis=2)
    j = d.argmin(1)
    ok = d[np.arange(len(ko)), j] < tol_kpt
    assert len(set(j[ok].tolist())) == int(ok.sum()), "two keypoints paired with the same oracle keypoint"
    un_out = [tuple(k) for k in ko[~ok]]
    un_ref = [tuple(kr[i]) for i in sorted(set(range(len(kr))) - set(j[ok].tolist()))]
    diff = un_out + un_ref
    if diff:
        assert ref_score_map is not None, f"{len(diff)} keypoints unpaired"
        assert len(diff) <= max_boundary, f"{len(diff)} keypoints unpaired"
        H, W = ref_score_map.shape
        for (x, y) in diff:
            x0, y0 = int(round(x)), int(round(y))
            win = ref_score_map[max(y0 - radius - 1, 0):min(y0 + radius + 2, H), max(x0 - radius - 1, 0):min(x0 + radius + 2, W)]
            near_thr = np.abs(win - threshold).min() < tol
            if not near_thr:  # n_limit cut: some pixel of the window ties with the smallest selected peak score
                peaks = [ref_score_map[int(round(py)), int(round(px))] for px, py in kr]
                assert np.abs(win - min(peaks)).min() < 10 * tol, f"keypoint {(x, y)} unpaired away from threshold / cut"
    a = np.nonzero(ok)[0]
    b = j[ok]
    dk = float(d[a, b].max()) if len(a) else 0.0
    ds = float(np.abs(out["scores"][a] - ref["scores"][b]).max()) if len(a) else 0.0
    dd = float(np.abs(out["descriptors"][:, a] - ref["descriptors"][:, b]).max()) if len(a) else 0.0
    assert ds < tol, f"score error {ds}"
    assert dd < tol, f"descriptor error {dd}"
    if len(ko) == len(kr) and not diff:  # same order as the reference (row-major, or score-descending under n_limit)
        order_same = bool(np.all(b == np.arange(len(b))))
    else:
        order_same = False
    return {"n": len(kr), "boundary_diffs": diff, "max_dkpt": dk, "max_dscore": ds, "max_ddesc": dd, "order_same": order_same}
