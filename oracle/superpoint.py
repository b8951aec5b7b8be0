"""SuperPoint CPU oracle (test infrastructure only - see oracle/__init__.py).

Restates, in functional torch-fp32 form, the graph the reference runs for
``SuperPointExtractor._extract``:

* adapter   : src/deep_image_matching/extractors/superpoint.py:107-146
* model     : thirdparty/SuperGluePretrainedNetwork/models/superpoint.py:160-227
* nms       : same file :47-63        (simple_nms, exact float equality)
* borders   : same file :66-71        (remove_borders)
* top-k     : same file :74-78        (top_k_keypoints)
* sampling  : same file :81-98        (mode "orig", align_corners=True)
              extractors/superpoint.py:16-27 (mode "fix", align_corners=False)

Weights are a dict name -> np.float32 array in PyTorch OIHW layout
(SURVEY Appendix D).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

ENC = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b"]
POOL_AFTER = {"conv1b", "conv2b", "conv3b"}

DEFAULT_CONF = {  # src/deep_image_matching/config.py:93-99 (superpoint+lightglue)
    "nms_radius": 3,
    "keypoint_threshold": 0.0005,
    "max_keypoints": 2048,
    "remove_borders": 4,
    "fix_sampling": False,
}


def load_weights_npz(path) -> dict:
    z = np.load(path)
    return {k: z[k].astype(np.float32) for k in z.files}


def _conv(x, w, name, pad):
    return F.conv2d(x, torch.from_numpy(w[name + ".weight"]), torch.from_numpy(w[name + ".bias"]), padding=pad)


def encoder(img01: torch.Tensor, w: dict) -> torch.Tensor:
    """img01: (1,1,H,W) in [0,1] -> (1,128,H/8,W/8). superpoint.py:161-171."""
    x = img01
    for name in ENC:
        x = F.relu(_conv(x, w, name, 1))
        if name in POOL_AFTER:
            x = F.max_pool2d(x, 2, 2)
    return x


def score_map(feat: torch.Tensor, w: dict) -> torch.Tensor:
    """(1,128,h,w) -> dense scores (8h, 8w). superpoint.py:174-179."""
    cpa = F.relu(_conv(feat, w, "convPa", 1))
    s = _conv(cpa, w, "convPb", 0)
    s = F.softmax(s, 1)[:, :-1]
    b, _, h, wd = s.shape
    s = s.permute(0, 2, 3, 1).reshape(b, h, wd, 8, 8)
    s = s.permute(0, 1, 3, 2, 4).reshape(b, h * 8, wd * 8)
    return s[0]


def simple_nms(scores: torch.Tensor, r: int) -> torch.Tensor:
    """superpoint.py:47-63 - two suppression refinement rounds, exact equality."""
    assert r >= 0
    s = scores[None, None]

    def mp(x):
        return F.max_pool2d(x, kernel_size=2 * r + 1, stride=1, padding=r)

    zeros = torch.zeros_like(s)
    max_mask = s == mp(s)
    for _ in range(2):
        supp = mp(max_mask.float()) > 0
        supp_s = torch.where(supp, zeros, s)
        new_max = supp_s == mp(supp_s)
        max_mask = max_mask | (new_max & (~supp))
    return torch.where(max_mask, s, zeros)[0, 0]


def select_keypoints(nms: torch.Tensor, conf: dict):
    """superpoint.py:183-210: threshold (row-major nonzero), border, top-k, flip to (x,y)."""
    H, W = nms.shape
    kp = torch.nonzero(nms > conf["keypoint_threshold"])  # (K,2) as (y,x), row-major
    sc = nms[kp[:, 0], kp[:, 1]]
    b = conf["remove_borders"]
    m = (kp[:, 0] >= b) & (kp[:, 0] < H - b) & (kp[:, 1] >= b) & (kp[:, 1] < W - b)
    kp, sc = kp[m], sc[m]
    k = conf["max_keypoints"]
    if k >= 0 and k < len(kp):
        sc, idx = torch.topk(sc, k, dim=0)
        kp = kp[idx]
    return torch.flip(kp, [1]).float(), sc


def dense_descriptors(feat: torch.Tensor, w: dict) -> torch.Tensor:
    """superpoint.py:213-216 -> (1,256,h,w) L2-normalised over channels."""
    cda = F.relu(_conv(feat, w, "convDa", 1))
    d = _conv(cda, w, "convDb", 0)
    return F.normalize(d, p=2, dim=1)


def sample_descriptors(kpts_xy: torch.Tensor, dense: torch.Tensor, fix_sampling: bool, s: int = 8):
    """Both variants; returns (256, N)."""
    b, c, h, w = dense.shape
    k = kpts_xy[None].clone()
    if fix_sampling:  # extractors/superpoint.py:16-27
        k = (k + 0.5) / (k.new_tensor([w, h]) * s)
        k = k * 2 - 1
        d = F.grid_sample(dense, k.view(b, 1, -1, 2), mode="bilinear", align_corners=False)
    else:  # thirdparty superpoint.py:81-98
        k = k - s / 2 + 0.5
        k = k / torch.tensor([(w * s - s / 2 - 0.5), (h * s - s / 2 - 0.5)]).to(k)[None]
        k = k * 2 - 1
        d = F.grid_sample(dense, k.view(b, 1, -1, 2), mode="bilinear", align_corners=True)
    d = F.normalize(d.reshape(b, c, -1), p=2, dim=1)
    return d[0]


@torch.no_grad()
def extract(image: np.ndarray, w: dict, conf: dict | None = None, return_debug: bool = False) -> dict:
    """Oracle of ``SuperPointExtractor._extract``.

    image: float32 (H,W) gray 0..255.  Returns the reference FeaturesDict:
    keypoints float32 (N,2) xy, scores float32 (N,), descriptors float32 (256,N).
    """
    conf = {**DEFAULT_CONF, **(conf or {})}
    assert image.ndim == 2
    # _frame2tensor: torch.tensor(image / 255.0, dtype=torch.float)   (:134-146)
    x = torch.tensor(image / 255.0, dtype=torch.float)[None, None]
    feat = encoder(x, w)
    dense_scores = score_map(feat, w)
    nms = simple_nms(dense_scores, conf["nms_radius"])
    kpts, scores = select_keypoints(nms, conf)
    dense = dense_descriptors(feat, w)
    desc = sample_descriptors(kpts, dense, conf["fix_sampling"])
    out = {
        "keypoints": kpts.numpy().copy(),
        "scores": scores.numpy().copy(),
        "descriptors": desc.numpy().copy(),
    }
    if return_debug:
        out["_feat"] = feat.numpy()
        out["_dense_scores"] = dense_scores.numpy()
        out["_nms"] = nms.numpy()
        out["_dense_desc"] = dense.numpy()
    return out


def canonical_order(feats: dict) -> np.ndarray:
    """Permutation sorting keypoints by (score desc, y asc, x asc) - removes topk tie ambiguity."""
    k = feats["keypoints"]
    return np.lexsort((k[:, 0], k[:, 1], -feats["scores"].astype(np.float64)))
