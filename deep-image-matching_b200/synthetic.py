"""Synthetic inputs of record for BASELINE config 2 ("blocks-16").

uint8 RGB images made of 16x16 random colour blocks, lightly blurred, give
~5-6k SuperPoint candidates at 1024x1024 (threshold 0.0005, nms 3) so the
top-2048 branch is exercised; image 1 of a pair is image 0 under a seeded mild
homography so LightGlue has true correspondences.
"""
from __future__ import annotations

import cv2
import numpy as np


def blocks_image(seed: int, size: int = 1024, block: int = 16) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n = size // block
    small = rng.integers(0, 256, (n, n, 3), dtype=np.uint8)
    img = cv2.resize(small, (size, size), interpolation=cv2.INTER_NEAREST)
    return cv2.GaussianBlur(img, (0, 0), 0.8)


def warp_pair(img0: np.ndarray, seed: int, jitter: float = 64.0) -> np.ndarray:
    rng = np.random.default_rng(seed + 7919)
    h, w = img0.shape[:2]
    src = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], np.float32)
    dst = src + rng.uniform(-jitter, jitter, (4, 2)).astype(np.float32)
    H = cv2.getPerspectiveTransform(src, dst)
    return cv2.warpPerspective(img0, H, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)


def to_gray_like_reference(rgb: np.ndarray) -> np.ndarray:
    """ExtractorBase.extract (extractor_base.py:190-202): the RGB array read by rasterio is
    passed to cv2.COLOR_BGR2GRAY (R/B weights swapped), then cast to float32."""
    return cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY).astype(np.float32)


def synthetic_pair(pair_id: int, size: int = 1024):
    """Returns (gray0, gray1) float32 (size,size) 0..255 as SuperPointExtractor._extract receives them."""
    a = blocks_image(pair_id, size)
    b = warp_pair(a, pair_id)
    return to_gray_like_reference(a), to_gray_like_reference(b)
