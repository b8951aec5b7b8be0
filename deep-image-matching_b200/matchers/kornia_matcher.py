"""KorniaMatcher on libdimb200 - drop-in for the reference plugin
(src/deep_image_matching/matchers/kornia_matcher.py:9-54): brute-force descriptor matching with kornia's
``DescriptorMatcher`` modes nn / mnn / snn / smnn, returning only the index pairs (:46-49).
"""
from __future__ import annotations

import numpy as np

from .. import _native
from ..config import Config
from .matcher_base import MatcherBase


class KorniaMatcher(MatcherBase):
    _default_conf = {"name": "kornia_matcher", "match_mode": "smnn", "th": 0.8}
    required_inputs = []
    min_matches = 20
    max_feat_no_tiling = 200000

    def __init__(self, config: Config) -> None:
        super().__init__(config)
        cfg = {**self._default_conf, **self.config.get("matcher", {})}
        if cfg["match_mode"] not in _native.NN_MODES:
            raise NotImplementedError(f"{cfg['match_mode']} is not supported. Try one of {list(_native.NN_MODES)}")
        self._mode, self._th = cfg["match_mode"], float(cfg["th"])
        self._ctx = _native.Context.get(int(self.config["general"].get("device", 0)))

    def _match_pairs(self, feats0: dict, feats1: dict) -> np.ndarray:
        # descriptors arrive (D,N); the reference transposes them to (N,D) for kornia (:36-37)
        idx, _ = self._ctx.nn_match(feats0["descriptors"], feats1["descriptors"], self._mode, self._th)
        return idx
