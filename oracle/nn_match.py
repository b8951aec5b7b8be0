"""Brute-force descriptor matcher CPU oracle (test infrastructure only).

PARITY UNPINNED for the kornia modes: ``KorniaMatcher._match_pairs``
(src/deep_image_matching/matchers/kornia_matcher.py:27-54) delegates to
``kornia.feature.DescriptorMatcher(match_mode, th)``; kornia (pinned 0.8.1 in
the reference's uv.lock) is neither vendored nor installable here, so the
functions below restate kornia 0.8.1's published ``match_nn / match_mnn /
match_snn / match_smnn`` (kornia/feature/matching.py) on top of ``torch.cdist``.
The reference keeps only the index array (kornia_matcher.py:46-49).

``hloc_mutual_nn`` restates the in-tree cosine mutual-NN matcher
(thirdparty/hloc/matchers/nearest_neighbor.py:6-56) and is pinned against it
by oracle/gen_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch


def _cdist(d1: torch.Tensor, d2: torch.Tensor) -> torch.Tensor:
    return torch.cdist(d1, d2)


def _no_match():
    return torch.empty(0, 1), torch.empty(0, 2, dtype=torch.long)


def match_nn(desc1, desc2, dm=None):
    if len(desc1) == 0 or len(desc2) == 0:
        return _no_match()
    dm = _cdist(desc1, desc2) if dm is None else dm
    dists, idx2 = torch.min(dm, dim=1)
    idx1 = torch.arange(0, idx2.size(0))
    return dists.view(-1, 1), torch.stack([idx1, idx2], 1)


def match_mnn(desc1, desc2, dm=None):
    if len(desc1) == 0 or len(desc2) == 0:
        return _no_match()
    dm = _cdist(desc1, desc2) if dm is None else dm
    ms = min(dm.size(0), dm.size(1))
    d12, idx2 = torch.min(dm, dim=1)
    d21, idx1 = torch.min(dm, dim=0)
    ar = torch.arange(ms)
    if dm.size(0) <= dm.size(1):
        mutual = ar == idx1[idx2][:ms]
        idxs = torch.stack([ar, idx2], 1)[mutual]
        dists = d12[mutual]
    else:
        mutual = ar == idx2[idx1][:ms]
        idxs = torch.stack([idx1, ar], 1)[mutual]
        dists = d21[mutual]
    return dists.view(-1, 1), idxs.view(-1, 2)


def match_snn(desc1, desc2, th=0.8, dm=None):
    if desc2.shape[0] < 2 or desc1.shape[0] == 0:
        return _no_match()
    dm = _cdist(desc1, desc2) if dm is None else dm
    vals, idx2 = torch.topk(dm, 2, dim=1, largest=False)
    ratio = vals[:, 0] / vals[:, 1]
    mask = ratio <= th
    dists = ratio[mask]
    if len(dists) == 0:
        return _no_match()
    idx1 = torch.arange(0, idx2.size(0))[mask]
    return dists.view(-1, 1), torch.stack([idx1, idx2[:, 0][mask]], 1)


def match_smnn(desc1, desc2, th=0.95, dm=None):
    if desc1.shape[0] < 2 or desc2.shape[0] < 2:
        return _no_match()
    dm = _cdist(desc1, desc2) if dm is None else dm
    dists1, idx1 = match_snn(desc1, desc2, th, dm)
    dists2, idx2 = match_snn(desc2, desc1, th, dm.t())
    if len(dists2) == 0 or len(dists1) == 0:
        return _no_match()
    idx2 = idx2.flip(1)
    # kornia intersects the two index lists with an L1 cdist; a set intersection is equivalent
    key1 = idx1[:, 0] * (dm.size(1) + 1) + idx1[:, 1]
    key2 = idx2[:, 0] * (dm.size(1) + 1) + idx2[:, 1]
    mutual1 = torch.isin(key1, key2)
    mutual2 = torch.isin(key2, key1)
    good1, good2 = idx1[mutual1], idx2[mutual2]
    dg1, dg2 = dists1[mutual1], dists2[mutual2]
    o1 = torch.sort(good1[:, 0]).indices
    o2 = torch.sort(good2[:, 0]).indices
    good1 = good1[o1]
    dists = torch.max(dg1[o1], dg2[o2])
    return dists.view(-1, 1), good1.view(-1, 2)


MODES = {"nn": match_nn, "mnn": match_mnn, "snn": match_snn, "smnn": match_smnn}


@torch.no_grad()
def kornia_match(feats0: dict, feats1: dict, match_mode: str = "smnn", th: float = 0.8):
    """Oracle of KorniaMatcher._match_pairs: returns (int64 (S,2) indices, float32 (S,) dist/ratio)."""
    d1 = torch.tensor(np.asarray(feats0["descriptors"]).T, dtype=torch.float)
    d2 = torch.tensor(np.asarray(feats1["descriptors"]).T, dtype=torch.float)
    if match_mode in ("snn", "smnn"):
        dist, idx = MODES[match_mode](d1, d2, th)
    else:
        dist, idx = MODES[match_mode](d1, d2)
    return idx.numpy().astype(np.int64).reshape(-1, 2), dist.numpy().astype(np.float32).reshape(-1)


@torch.no_grad()
def hloc_mutual_nn(desc0: np.ndarray, desc1: np.ndarray, ratio_thresh=None, distance_thresh=None, mutual=True):
    """hloc NearestNeighbor on (D,N),(D,M) unit descriptors -> matches0 (N,) int64 (-1 = none), scores0."""
    a, b = torch.tensor(desc0, dtype=torch.float), torch.tensor(desc1, dtype=torch.float)
    if a.shape[-1] == 0 or b.shape[-1] == 0:
        return np.full(a.shape[-1], -1, np.int64), np.zeros(a.shape[-1], np.float32)
    if a.shape[-1] == 1 or b.shape[-1] == 1:
        ratio_thresh = None
    sim = a.t() @ b

    def find(s):
        v, ind = s.topk(2 if ratio_thresh else 1, dim=-1, largest=True)
        dist = 2 * (1 - v)
        mask = torch.ones(ind.shape[:-1], dtype=torch.bool)
        if ratio_thresh:
            mask &= dist[..., 0] <= (ratio_thresh**2) * dist[..., 1]
        if distance_thresh:
            mask &= dist[..., 0] <= distance_thresh**2
        return torch.where(mask, ind[..., 0], ind.new_tensor(-1)), torch.where(mask, (v[..., 0] + 1) / 2, v.new_tensor(0))

    m0, s0 = find(sim)
    if mutual:
        m1, _ = find(sim.t())
        loop = torch.gather(m1, -1, torch.where(m0 > -1, m0, m0.new_tensor(0)))
        m0 = torch.where((m0 > -1) & (torch.arange(m0.shape[-1]) == loop), m0, m0.new_tensor(-1))
    return m0.numpy().astype(np.int64), s0.numpy().astype(np.float32)
