"""features.h5 boundary of the hot path.

Mirrors ``save_features_h5`` (reference extractors/extractor_base.py:56-99: every ndarray is stored as
float16, gzip level 9, one group per image) and ``get_features`` (io/h5.py:45-89: keypoints/descriptors/
scores/tile_idx back to float32, image_size to int32).  h5py is optional in this environment: without it the
same arrays (after the same float16 cast) go to an ``.npz`` per image inside ``<path>.d/`` so the hot path
still sees exactly the values the reference would read back.
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np

try:  # pragma: no cover - depends on the environment
    import h5py
except ImportError:  # pragma: no cover
    h5py = None


def as_half_roundtrip(features: dict) -> dict:
    """The value-level effect of save_features_h5 + get_features: everything passes through fp16."""
    out = {}
    for k, v in features.items():
        if not isinstance(v, np.ndarray):
            continue
        h = v.astype(np.float16)
        out[k] = h.astype(np.int32) if k == "image_size" else h.astype(np.float32)
    return out


def save_features_h5(feature_path, features: dict, im_name: str, as_half: bool = True):
    feat_dtype = np.float16 if as_half else np.float32
    for k, v in features.items():
        if not isinstance(v, np.ndarray):
            raise TypeError(f"Features data must be of type np.ndarray, not {type(v)}")
    if h5py is not None:
        with h5py.File(str(feature_path), "a", libver="latest") as fd:
            if im_name in fd:
                del fd[im_name]
            grp = fd.create_group(im_name)
            for k, v in features.items():
                grp.create_dataset(k, data=v, dtype=feat_dtype, compression="gzip", compression_opts=9)
        return
    d = Path(str(feature_path) + ".d")
    d.mkdir(parents=True, exist_ok=True)
    np.savez(d / (im_name + ".npz"), **{k: v.astype(feat_dtype) for k, v in features.items()})


def get_features(path, name: str) -> dict:
    if h5py is not None and os.path.exists(str(path)):
        with h5py.File(str(path), "r", libver="latest") as fd:
            if name not in fd:
                raise ValueError(f"Cannot find image {name} in {path}")
            raw = {k: np.array(fd[name][k]) for k in fd[name]}
    else:
        f = Path(str(path) + ".d") / (name + ".npz")
        if not f.exists():
            raise ValueError(f"Cannot find image {name} in {path}")
        z = np.load(f)
        raw = {k: z[k] for k in z.files}
    if "keypoints" not in raw or "descriptors" not in raw:
        raise KeyError(f"Cannot find keypoints and descriptors in {path}")
    feats = {"keypoints": raw["keypoints"].astype(np.float32), "descriptors": raw["descriptors"].astype(np.float32)}
    for k in ("tile_idx", "scores"):
        if k in raw:
            feats[k] = raw[k].astype(np.float32)
    if "image_size" in raw:
        feats["image_size"] = raw["image_size"].astype(np.int32)
    return feats


class FeatureStore:
    """In-memory replacement of the per-pair features.h5 round trip.

    The reference gzip-9-writes every image's features as float16 (``save_features_h5``, extractor_base.py:56-99) and re-opens
    and re-reads the file for both images of every pair (``get_features``, io/h5.py:45-89, from matcher_base.py:221-222).  This
    store holds the same float16 content in device memory (``dimb_fstore``, csrc/fstore.cu), keyed by image name:

    * ``put`` / ``put_dev`` apply the writer's float16 cast once;
    * ``get_features(name)`` honours the reader's contract - float32 ``keypoints (N,2)``, ``descriptors (D,N)``, ``scores``,
      ``tile_idx`` whose values are float16-exact, ``image_size`` int32 - and raises ``ValueError`` for an unknown image like the
      reader does (io/h5.py:60-61);
    * ``feats_dev(name)`` hands the block to the device matchers without any copy;
    * ``write_h5(path)`` emits the whole store in one pass at the end (one bulk device->host copy per image; needs h5py for a
      real HDF5 file, otherwise the ``<path>.d/`` mirror of this module).
    """

    def __init__(self, ctx, max_images: int, cap: int, desc_dim: int):
        from . import _native
        self.dev = _native.FeatureStoreDev(ctx, max_images, cap, desc_dim)
        self.names: dict = {}

    def slot(self, name: str, create: bool = False) -> int:
        if name not in self.names:
            if not create:
                raise ValueError(f"Cannot find image {name} in the feature store")
            if len(self.names) >= self.dev.n_slots:
                raise RuntimeError(f"feature store is full ({self.dev.n_slots} images)")
            self.names[name] = len(self.names)
        return self.names[name]

    def put(self, name: str, features: dict):
        for k, v in features.items():
            if not isinstance(v, np.ndarray):
                raise TypeError(f"Features data must be of type np.ndarray, not {type(v)}")  # extractor_base.py:71-75
        self.dev.put(self.slot(name, True), features)

    def put_dev(self, name: str, d_kpts, d_scores, d_desc, desc_ld, d_count, height, width, stream=0):
        self.dev.put_dev(self.slot(name, True), d_kpts, d_scores, d_desc, desc_ld, d_count, height, width, None, stream)

    def get_features(self, name: str) -> dict:
        return self.dev.get(self.slot(name))

    def feats_dev(self, name: str):
        return self.dev.feats_dev(self.slot(name))

    def write_h5(self, path):
        for name in self.names:
            save_features_h5(path, self.get_features(name), name, as_half=True)
