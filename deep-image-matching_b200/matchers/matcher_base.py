"""Mirror of the reference's matcher plugin surface (src/deep_image_matching/matchers/matcher_base.py:63-340).

Restated: constructor contract (:97-141), abstract ``_match_pairs`` (:164-183) and the no-tiling branch of
``match`` (:185-340) up to raw_matches.  Tile selection and geometric verification are out of scope
(SURVEY 2.1); they run unchanged on top of ``_match_pairs`` in the reference.
"""
from __future__ import annotations

import inspect
from abc import ABCMeta, abstractmethod
from pathlib import Path

import numpy as np

from ..config import Config
from ..io_h5 import get_features


def matcher_loader(root, model):
    """matcher_base.py:36-60."""
    module_path = f"{root.__name__}.{model}"
    module = __import__(module_path, fromlist=[""])
    classes = inspect.getmembers(module, inspect.isclass)
    classes = [c for c in classes if c[1].__module__ == module_path]
    classes = [c for c in classes if issubclass(c[1], MatcherBase)]
    assert len(classes) == 1, classes
    return classes[0][1]


class MatcherBase(metaclass=ABCMeta):
    _default_general_conf = {"force_cpu": False, "min_inliers_per_pair": 15, "min_inlier_ratio_per_pair": 0.2}
    _default_conf = {}
    required_inputs = []
    min_matches = 20
    max_feat_no_tiling = 20000

    def __init__(self, custom_config: Config) -> None:
        if not isinstance(custom_config, Config):
            raise TypeError("Invalid config object. 'custom_config' must be a Config object")
        self.config = {
            "general": {**self._default_general_conf, **custom_config.general},
            "matcher": {**self._default_conf, **custom_config.matcher},
        }
        if self.config["general"].get("force_cpu"):
            raise RuntimeError("dim_b200 has no CPU path (force_cpu=True is not supported)")
        self._device = "cuda"

    @abstractmethod
    def _match_pairs(self, feats0: dict, feats1: dict) -> np.ndarray:
        raise NotImplementedError("Subclasses must implement _match_pairs() method.")

    def _match_by_tile(self, features0: dict, features1: dict, tile_pairs, select_unique: bool = True) -> np.ndarray:
        """matcher_base.py:362-485 after tile selection (``tiling.tile_selection``): ``_match_pairs`` on the features of each
        selected tile pair (each keeps the full-image ``image_size``), indices mapped back to the full arrays, ``np.unique``."""
        from ..tiling import get_features_by_tile

        matches_full = np.zeros((0, 2), np.int64)
        for tidx0, tidx1 in tile_pairs:
            f0, idx0 = get_features_by_tile(features0, tidx0)
            f1, idx1 = get_features_by_tile(features1, tidx1)
            corr = self._match_pairs(f0, f1)
            orig = np.zeros_like(corr)
            orig[:, 0], orig[:, 1] = idx0[corr[:, 0]], idx1[corr[:, 1]]
            matches_full = np.vstack((matches_full, orig))
        if select_unique and len(matches_full):
            matches_full = np.unique(matches_full, axis=0)
        return matches_full

    def match(self, feature_path: Path, matches_path: Path, img0: Path, img1: Path) -> np.ndarray:
        """No-tiling branch of MatcherBase.match up to the raw matches (:218-296)."""
        img0_name, img1_name = Path(img0).name, Path(img1).name
        features0 = get_features(feature_path, img0_name)
        features1 = get_features(feature_path, img1_name)
        n = max(len(features0["keypoints"]), len(features1["keypoints"]))
        if n > self.max_feat_no_tiling:
            raise RuntimeError("CUDA out of memory. too many features for full-image matching, use tiling")  # :229-236
        matches = self._match_pairs(features0, features1)
        if matches.shape[0] < self.min_matches:
            return None  # :292-296 (reference logs and skips the pair)
        return matches
