"""SuperGlueMatcher on libdimb200 - drop-in for the reference plugin
(src/deep_image_matching/matchers/superglue.py:54-106): same class name, attributes and
``_match_pairs(feats0, feats1) -> int64 (S,2)`` contract (``correspondence_matrix_from_matches0``: rows ascending in index 0).

Reproduced quirk: the reference builds the model config from ``self._default_conf`` (MatcherBase's, empty) instead of its own
``default_config`` (:55-60,72), so ``sinkhorn_iterations 20 / match_threshold 0.3`` never reach the model - SuperGlue's own
defaults (100 / 0.2, thirdparty/SuperGluePretrainedNetwork/models/superglue.py:213-220) apply unless the user's matcher config
sets them.  ``scores`` and ``image_size`` are required in the features (features_2_sg :8-41).
"""
from __future__ import annotations

import os

import numpy as np

from .. import _native
from ..config import Config
from ..weights import from_torch_checkpoint, load_npz
from .matcher_base import MatcherBase

SUPERGLUE_DEFAULTS = {"weights": "outdoor", "sinkhorn_iterations": 100, "match_threshold": 0.2, "GNN_layers": ["self", "cross"] * 9}


class SuperGlueMatcher(MatcherBase):
    default_config = {  # sic: unused by the reference (see module docstring)
        "name": "superglue",
        "weights": "outdoor",
        "sinkhorn_iterations": 20,
        "match_threshold": 0.3,
    }
    required_inputs = []
    min_matches = 20
    max_feat_no_tiling = 50000

    def __init__(self, config: Config) -> None:
        super().__init__(config)
        self._cfg = {**SUPERGLUE_DEFAULTS, **self._default_conf, **self.config.get("matcher", {})}
        if self._cfg["weights"] not in ("indoor", "outdoor"):
            raise AssertionError(self._cfg["weights"])  # superglue.py:243
        self._ctx = _native.Context.get(int(self.config["general"].get("device", 0)))
        w = self._cfg.get("weights_dict")
        if w is None:
            path = self._cfg.get("weights_path") or os.environ.get("DIMB_SUPERGLUE_WEIGHTS")
            if path is None:
                raise FileNotFoundError(f"superglue_{self._cfg['weights']}.pth: pass matcher['weights_path'] or set DIMB_SUPERGLUE_WEIGHTS "
                                        "(the reference reads it from thirdparty/SuperGluePretrainedNetwork/models/weights/)")
            w = load_npz(path) if str(path).endswith(".npz") else from_torch_checkpoint(path)
        self._weights = w
        self._net = None
        self._cap = 0

    def _ensure(self, kpts):
        if self._net is None or kpts > self._cap:
            self._cap = max(kpts, self._cap, 2048)
            c = self._cfg
            self._net = _native.SuperGlueNet(self._ctx, self._weights, gnn_layers=tuple(c["GNN_layers"]),
                                             sinkhorn_iterations=c["sinkhorn_iterations"], match_threshold=c["match_threshold"],
                                             max_kpts=self._cap)
        return self._net

    def _match_pairs(self, feats0: dict, feats1: dict) -> np.ndarray:
        for f in (feats0, feats1):
            for k in ("keypoints", "descriptors", "scores", "image_size"):
                if k not in f:
                    raise KeyError(f"SuperGlue needs '{k}' in the features")
        kmax = max(feats0["keypoints"].shape[0], feats1["keypoints"].shape[0])
        return self._ensure(kmax).match(feats0, feats1)["matches"]
