"""CPU oracle of the tiled flows around the hot path (test infrastructure only; nothing in the product imports it).

Restates, on top of the oracle networks (oracle/superpoint.py, oracle/aliked.py, oracle/lightglue.py):
  * Tiler.compute_tiles_by_size            reference utils/tiling.py:63-192 (kornia compute_padding without stride, quirk A.7)
  * ExtractorBase._extract_by_tile         reference extractors/extractor_base.py:279-390
  * tile_selection (EXHAUSTIVE/GRID/PRESELECTION)  reference matchers/matcher_base.py:989-1148, preselection nets :143-159
  * MatcherBase._match_by_tile             reference matchers/matcher_base.py:362-485, get_features_by_tile :1380-1391
Pinned against the reference's own tiling unit tests (tests/test_tiling.py:22-157: tile counts, padding) in
tests/test_host_logic.py; the networks are pinned by tests/golden (see the oracle modules).
"""
from __future__ import annotations

from itertools import product

import numpy as np


def compute_tiles(image: np.ndarray, window_xy, overlap):
    wh = (window_xy, window_xy) if isinstance(window_xy, int) else (window_xy[1], window_xy[0])  # -> (H, W)
    ov = (overlap, overlap) if isinstance(overlap, int) else (overlap[1], overlap[0])
    arr = image if image.ndim == 3 else image[:, :, None]
    H, W = arr.shape[:2]
    pads = []
    for size, win in ((H, wh[0]), (W, wh[1])):  # kornia.contrib.compute_padding with stride = window
        rem = (size - win) % win
        p = win - rem if rem else 0
        pads.append((p // 2, p - p // 2))
    padded = np.pad(arr, (pads[0], pads[1], (0, 0)))
    sy, sx = wh[0] - ov[0], wh[1] - ov[1]
    ys = list(range(0, padded.shape[0] - wh[0] + 1, sy))
    xs = list(range(0, padded.shape[1] - wh[1] + 1, sx))
    tiles, origins = {}, {}
    for r, y in enumerate(ys):
        for c, x in enumerate(xs):
            k = r * len(xs) + c
            tiles[k] = padded[y:y + wh[0], x:x + wh[1]]
            origins[k] = (x - pads[1][0], y - pads[0][0])
    return tiles, origins, (pads[0][0], pads[0][1], pads[1][0], pads[1][1])


def extract_by_tile(image: np.ndarray, extract_fn, tile_size, overlap, descriptor_size: int):
    """extract_fn(tile (h,w[,C])) -> dict(keypoints (N,2), descriptors (D,N), scores (N,))."""
    tiles, origins, _ = compute_tiles(image, tile_size, overlap)
    K, D, S, T = [], [], [], []
    for idx, tile in tiles.items():
        f = extract_fn(tile)
        kp = f["keypoints"] + np.array(origins[idx], dtype=f["keypoints"].dtype)
        m = (kp[:, 0] >= 2) & (kp[:, 0] < image.shape[1] - 2) & (kp[:, 1] >= 2) & (kp[:, 1] < image.shape[0] - 2)
        if m.sum() > 0:
            K.append(kp[m]); D.append(f["descriptors"][:, m]); S.append(f["scores"][m]); T.append(np.full(int(m.sum()), idx, np.float32))
    if not K:
        return {"keypoints": np.zeros((0, 2), np.float32), "descriptors": np.zeros((descriptor_size, 0), np.float32),
                "scores": np.zeros(0, np.float32), "tile_idx": np.zeros(0, np.float32)}
    kp, de, sc, ti = np.vstack(K), np.hstack(D), np.concatenate(S), np.concatenate(T)
    kp, u = np.unique(kp, axis=0, return_index=True)
    return {"keypoints": kp, "descriptors": de[:, u], "scores": sc[u], "tile_idx": ti[u]}


def features_by_tile(features: dict, t: int):
    sel = features["tile_idx"] == t
    return {"keypoints": features["keypoints"][sel], "descriptors": features["descriptors"][:, sel], "scores": features["scores"][sel],
            "image_size": features["image_size"]}, np.where(sel)[0]


def match_by_tile(features0: dict, features1: dict, tile_pairs, match_fn):
    """match_fn(feats0, feats1) -> int64 (S,2)."""
    full = np.zeros((0, 2), np.int64)
    for t0, t1 in tile_pairs:
        f0, i0 = features_by_tile(features0, t0)
        f1, i1 = features_by_tile(features1, t1)
        c = match_fn(f0, f1)
        full = np.vstack((full, np.stack([i0[c[:, 0]], i1[c[:, 1]]], 1).astype(np.int64).reshape(-1, 2)))
    return np.unique(full, axis=0) if len(full) else full


def preselection_keypoints(i0, i1, size, sp_weights, lg_weights):
    """Down-sampled SuperPoint (hloc wrapper semantics: fix_sampling) + LightGlue without image_size (matcher_base.py:1054-1089)."""
    import cv2

    from . import lightglue as o_lg
    from . import superpoint as o_sp
    conf_sp = {"nms_radius": 5, "max_keypoints": 4000, "keypoint_threshold": 0.005, "fix_sampling": True}
    conf_lg = {**o_lg.DEFAULT_CONF, "depth_confidence": 0.9, "width_confidence": 0.95, "filter_threshold": 0.3}
    feats, scales = [], []
    for im in (i0, i1):
        wh = im.shape[:2][::-1]
        s = size / max(wh)
        low = cv2.resize(im, tuple(int(round(x * s)) for x in wh), interpolation=cv2.INTER_AREA)
        f = o_sp.extract(low, sp_weights, conf_sp)
        feats.append({"keypoints": f["keypoints"], "descriptors": f["descriptors"]})
        scales.append(s)
    m = o_lg.match(feats[0], feats[1], lg_weights, conf_lg)["matches"]
    return feats[0]["keypoints"][m[:, 0]] / scales[0], feats[1]["keypoints"][m[:, 1]] / scales[1]


def select_tiles(i0, i1, method, tile_size, overlap, kp0=None, kp1=None, min_matches_per_tile=5):
    t0, o0, _ = compute_tiles(i0, tile_size, overlap)
    t1, o1, _ = compute_tiles(i1, tile_size, overlap)
    if method == "exhaustive":
        return sorted(product(t0.keys(), t1.keys()))
    if method == "grid":
        return sorted(zip(t0.keys(), t1.keys()))
    out = set()
    for a, b in sorted(product(t0.keys(), t1.keys())):
        r0 = np.array([o0[a][0], o0[a][1], o0[a][0] + tile_size[0], o0[a][1] + tile_size[1]])
        r1 = np.array([o1[b][0], o1[b][1], o1[b][0] + tile_size[0], o1[b][1] + tile_size[1]])
        in0 = np.all(kp0 > r0[:2], 1) & np.all(kp0 < r0[2:], 1)
        in1 = np.all(kp1 > r1[:2], 1) & np.all(kp1 < r1[2:], 1)
        if int((in0 & in1).sum()) > min_matches_per_tile:
            out.add((a, b))
    return sorted(out)
