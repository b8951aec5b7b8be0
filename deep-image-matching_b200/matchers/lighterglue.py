"""LighterGlueMatcher on libdimb200 - drop-in for the reference plugin
(src/deep_image_matching/matchers/lighterglue.py:78-262): same class name, attributes, constructor
``(config, local_features="xfeat")`` and ``_match_pairs(feats0, feats1) -> int64 (S,2)`` contract.

LighterGlue (thirdparty/accelerated_features/modules/lighterglue.py:12-48) is the LightGlue architecture with
``input_dim 64, descriptor_dim 96, one head, 6 layers`` and its own trained checkpoint ``xfeat-lighterglue.pt``
(vendored by the reference).  Reproduced behaviour of the reference plugin: only ``filter_threshold`` of the plugin's
config reaches the network (``min_conf``, :243-244) - depth / width confidences are LighterGlue's own defaults (-1 / 0.95,
lighterglue.py:22-24), the plugin's 0.95 / 0.99 are never used; ``image_size`` is mandatory and is swapped from DIM's
``[H,W]`` to ``[W,H]`` before the call (:214-216), unlike LightGlueMatcher.
The shape is served by the shape-generic fp32 kernels of ``csrc/lightglue_generic.cu``.
"""
from __future__ import annotations

import os

import numpy as np

from .. import _native
from ..config import Config
from ..weights import DATA, load_npz
from .lightglue import featuresDict2Lightglue
from .matcher_base import MatcherBase

LIGHTERGLUE_CONF = {"input_dim": 64, "descriptor_dim": 96, "n_layers": 6, "num_heads": 1, "depth_confidence": -1,
                    "width_confidence": 0.95}  # modules/lighterglue.py:12-27


def lighterglue_weights(path=None) -> dict:
    """``matcher.*`` tensors of xfeat-lighterglue.pt with the key renames of modules/lighterglue.py:40-46."""
    for p in (path, os.environ.get("DIMB_LIGHTERGLUE_WEIGHTS"), os.path.join(DATA, "lighterglue_weights.npz")):
        if not p or not os.path.exists(p):
            continue
        if str(p).endswith(".npz"):
            return load_npz(p)
        import torch
        sd = {k: v for k, v in torch.load(str(p), map_location="cpu").items() if k.startswith("matcher.")}
        for i in range(LIGHTERGLUE_CONF["n_layers"]):
            sd = {k.replace(f"self_attn.{i}", f"transformers.{i}.self_attn"): v for k, v in sd.items()}
            sd = {k.replace(f"cross_attn.{i}", f"transformers.{i}.cross_attn"): v for k, v in sd.items()}
        return {k.replace("matcher.", ""): v.numpy().astype(np.float32) for k, v in sd.items() if v.dtype.is_floating_point}
    raise FileNotFoundError("xfeat-lighterglue weights not found (set DIMB_LIGHTERGLUE_WEIGHTS)")


class LighterGlueMatcher(MatcherBase):
    _default_conf = {
        "flash": True,
        "mp": False,
        "depth_confidence": 0.95,  # sic: never reaches the network (see module docstring)
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
        if self._localfeatures != "xfeat":
            raise ValueError(f"Unsupported local feature extractor: {self._localfeatures}")
        self._cfg = {**self._default_conf, **self.config.get("matcher", {})}
        self._ctx = _native.Context.get(int(self.config["general"].get("device", 0)))
        self._weights = self._cfg.get("weights_dict") or lighterglue_weights(self._cfg.get("weights"))
        self._net = None
        self._cap = 0

    def _ensure(self, kpts):
        if self._net is None or kpts > self._cap:
            self._cap = max(kpts, self._cap, 2048)
            self._net = _native.LightGlueNet(self._ctx, self._weights, filter_threshold=self._cfg["filter_threshold"], max_pairs=1,
                                             max_kpts=self._cap, **LIGHTERGLUE_CONF)
        return self._net

    def _match_pairs(self, feats0: dict, feats1: dict) -> np.ndarray:
        return self.match_scored(feats0, feats1)["matches"]

    def match_scored(self, feats0: dict, feats1: dict) -> dict:
        conv = []
        for f in (feats0, feats1):
            c = featuresDict2Lightglue(f)
            if "image_size" not in c:
                raise ValueError("image_size not found in features - required by XFeat's match_lighterglue")  # :192-193
            hw = np.asarray(c["image_size"], np.float32).ravel()
            c["image_size"] = np.array([hw[1], hw[0]], np.float32)  # [H,W] -> [W,H] (:214-216)
            conv.append(c)
        kmax = max(c["keypoints"].shape[0] for c in conv)
        return self._ensure(kmax).match([tuple(conv)])[0]
