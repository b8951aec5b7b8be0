"""LightGlueMatcher on libdimb200 - drop-in for the reference plugin
(src/deep_image_matching/matchers/lightglue.py:69-125): same class name, attributes, constructor
``(config, local_features=...)`` and ``_match_pairs(feats0, feats1) -> int64 (S,2)`` contract.
"""
from __future__ import annotations

import os

import numpy as np

from .. import _native
from ..config import Config
from ..weights import from_torch_checkpoint, load_npz
from .matcher_base import MatcherBase

# LightGlue.features (thirdparty/LightGlue/lightglue/lightglue.py:330-351)
FEATURES = {"superpoint": 256, "disk": 128, "aliked": 128, "sift": 128}


def featuresDict2Lightglue(feats: dict) -> dict:
    """matchers/lightglue.py:8-66 without the torch conversion: decide the descriptor layout from N."""
    feats = {k: v[0] if isinstance(v, (list, tuple)) else v for k, v in feats.items()}
    if "keypoints" not in feats or "descriptors" not in feats:
        raise KeyError("features must contain 'keypoints' and 'descriptors'")
    kpts, desc = np.asarray(feats["keypoints"]), np.asarray(feats["descriptors"])
    if kpts.ndim != 2 or kpts.shape[1] != 2:
        raise ValueError(f"Invalid keypoints shape: {kpts.shape}")
    n = kpts.shape[0]
    if desc.ndim != 2:
        raise ValueError(f"Invalid descriptors shape: {desc.shape}")
    if desc.shape[1] == n and desc.shape[0] != n:
        layout = 0  # (D,N): the library reads it transposed, no host copy
    elif desc.shape[0] == n:
        layout = 1
    else:
        raise ValueError(f"Descriptor / keypoint mismatch: descriptors={desc.shape}, keypoints={kpts.shape}")
    out = {"keypoints": kpts, "descriptors": desc, "_layout": layout}
    if feats.get("image_size") is not None:
        out["image_size"] = np.asarray(feats["image_size"], np.float32)
    return out


class LightGlueMatcher(MatcherBase):
    _default_conf = {
        "flash": True,
        "mp": False,
        "depth_confidence": 0.95,
        "width_confidence": 0.99,
        "filter_threshold": 0.1,
        "weights": None,
    }
    required_inputs = []
    min_matches = 20
    max_feat_no_tiling = 200000

    def __init__(self, config: Config, local_features="superpoint") -> None:
        self._localfeatures = local_features
        super().__init__(config)
        cfg = {**self._default_conf, **self.config.get("matcher", {})}
        self._cfg = cfg
        if cfg.get("mp"):
            raise RuntimeError("mixed precision (mp=True) is not implemented; use precision='fast' instead")
        if local_features not in FEATURES:
            print(f"Unsupported features: {local_features} not in {{{','.join(FEATURES)}}}")  # lightglue.py:354-356
        self._input_dim = cfg.get("input_dim", FEATURES.get(local_features, 256))
        self._n_layers = cfg.get("n_layers", 9)
        self._ctx = _native.Context.get(int(self.config["general"].get("device", 0)))
        self._weights = self._load_weights(cfg)
        self._net = None
        self._cap = (0, 0)
        if self._localfeatures == "disk":
            self.max_feat_no_tiling = 50000

    def _load_weights(self, cfg) -> dict:
        w = cfg.get("weights_dict")
        if w is not None:
            return w
        path = cfg.get("weights") or os.environ.get("DIMB_LIGHTGLUE_WEIGHTS")
        if path is None:
            raise FileNotFoundError(
                f"{self._localfeatures}_lightglue weights: the reference downloads them from GitHub releases "
                "(lightglue.py:381-384); offline, pass matcher['weights']=<.pth|.npz> or set DIMB_LIGHTGLUE_WEIGHTS")
        return load_npz(path) if str(path).endswith(".npz") else from_torch_checkpoint(path)

    def _ensure(self, pairs, kpts):
        p, k = self._cap
        if self._net is None or pairs > p or kpts > k:
            self._cap = (max(pairs, p), max(kpts, k, 2048))
            c = self._cfg
            self._net = _native.LightGlueNet(
                self._ctx, self._weights, input_dim=self._input_dim, descriptor_dim=256, n_layers=self._n_layers,
                num_heads=4, depth_confidence=c["depth_confidence"], width_confidence=c["width_confidence"],
                filter_threshold=c["filter_threshold"], prune_min_kpts=c.get("prune_min_kpts", 1536),
                max_pairs=self._cap[0], max_kpts=self._cap[1])
        return self._net

    def _match_pairs(self, feats0: dict, feats1: dict) -> np.ndarray:
        return self.match_many([(feats0, feats1)])[0]

    def match_many(self, pairs, return_scores: bool = False) -> list:
        """Batched entry (the reference is batch-1): list of (feats0, feats1) -> list of int64 (S,2)."""
        conv = [(featuresDict2Lightglue(a), featuresDict2Lightglue(b)) for a, b in pairs]
        kmax = max(max(a["keypoints"].shape[0], b["keypoints"].shape[0]) for a, b in conv)
        res = self._ensure(len(conv), kmax).match(conv)
        if return_scores:
            return res
        return [r["matches"] for r in res]
