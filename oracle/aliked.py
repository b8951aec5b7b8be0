"""ALIKED CPU oracle (test infrastructure only - see oracle/__init__.py).

Functional torch-fp32 restatement of the graph the reference runs for ``AlikedExtractor._extract``:

* adapter          : src/deep_image_matching/extractors/aliked.py:45-85
* model            : thirdparty/LightGlue/lightglue/aliked.py:560-693 (ALIKED), blocks :367-449,
                     DeformableConv2d :274-330 (torchvision.ops.deform_conv2d), InputPadder :247-271
* detector (DKD)   : :92-244, simple_nms :66-89
* descriptor (SDDH): :452-558, get_patches :48-63

Reproduced quirk: ``ALIKED.forward`` unpacks DKD's ``(keypoints, dispersity, scores)`` as
``(keypoints, kptscores, scoredispersitys)`` (:682), so the emitted ``keypoint_scores`` are the score *dispersities*.

Weights: dict name -> np.float32 array with the reference's state_dict names.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
import torchvision

CFGS = {  # c1, c2, c3, c4, dim, K, M   (aliked.py:574-579)
    "aliked-t16": [8, 16, 32, 64, 64, 3, 16],
    "aliked-n16": [16, 32, 64, 128, 128, 3, 16],
    "aliked-n16rot": [16, 32, 64, 128, 128, 3, 16],
    "aliked-n32": [16, 32, 64, 128, 128, 3, 32],
}
DEFAULT_CONF = {  # extractors/aliked.py:23-30 / config.py:198-203
    "model_name": "aliked-n16rot",
    "max_num_keypoints": 4000,
    "detection_threshold": 0.2,
    "nms_radius": 2,
}
N_LIMIT_MAX = 20000


def _t(w, k):
    return torch.from_numpy(w[k])


def _bn(x, w, p):
    return F.batch_norm(x, _t(w, p + ".running_mean"), _t(w, p + ".running_var"), _t(w, p + ".weight"), _t(w, p + ".bias"), False, 0.0, 1e-5)


def _conv(x, w, p, pad=1):
    b = w.get(p + ".bias")
    return F.conv2d(x, _t(w, p + ".weight"), None if b is None else torch.from_numpy(b), padding=pad)


def _dcn(x, w, p):
    """DeformableConv2d.forward (:305-330)."""
    h, wd = x.shape[2:]
    max_offset = max(h, wd) / 4.0
    offset = _conv(x, w, p + ".offset_conv").clamp(-max_offset, max_offset)
    b = w.get(p + ".regular_conv.bias")
    return torchvision.ops.deform_conv2d(input=x, offset=offset, weight=_t(w, p + ".regular_conv.weight"),
                                         bias=None if b is None else torch.from_numpy(b), padding=(1, 1), mask=None)


def _any_conv(x, w, p, dcn):
    return _dcn(x, w, p) if dcn else _conv(x, w, p)


def _resblock(x, w, p, dcn):
    """ResBlock.forward (:431-449) with the 1x1 downsample on the identity."""
    out = F.selu(_bn(_any_conv(x, w, p + ".conv1", dcn), w, p + ".bn1"))
    out = _bn(_any_conv(out, w, p + ".conv2", dcn), w, p + ".bn2")
    return F.selu(out + _conv(x, w, p + ".downsample", 0))


def simple_nms(scores, r):
    zeros = torch.zeros_like(scores)
    mp = lambda t: F.max_pool2d(t, kernel_size=r * 2 + 1, stride=1, padding=r)
    max_mask = scores == mp(scores)
    for _ in range(2):
        supp = mp(max_mask.float()) > 0
        ss = torch.where(supp, zeros, scores)
        new = ss == mp(ss)
        max_mask = max_mask | (new & (~supp))
    return torch.where(max_mask, scores, zeros)


def dense_maps(img01: torch.Tensor, w: dict, cfg):
    """extract_dense_map (:644-675): (1,3,H,W) in [0,1] -> feature_map (1,dim,H,W) L2-normalised, score_map (1,1,H,W)."""
    c1, c2, c3, c4, dim, K, M = cfg
    H, W = img01.shape[-2:]
    div = 32
    ph = (((H // div) + 1) * div - H) % div
    pw = (((W // div) + 1) * div - W) % div
    pad = [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]
    x = F.pad(img01, pad, mode="replicate")
    x1 = F.selu(_bn(_conv(x, w, "block1.conv1"), w, "block1.bn1"))
    x1 = F.selu(_bn(_conv(x1, w, "block1.conv2"), w, "block1.bn2"))
    x2 = _resblock(F.avg_pool2d(x1, 2, 2), w, "block2", False)
    x3 = _resblock(F.avg_pool2d(x2, 4, 4), w, "block3", True)
    x4 = _resblock(F.avg_pool2d(x3, 4, 4), w, "block4", True)
    x1 = F.selu(_conv(x1, w, "conv1", 0))
    x2 = F.selu(_conv(x2, w, "conv2", 0))
    x3 = F.selu(_conv(x3, w, "conv3", 0))
    x4 = F.selu(_conv(x4, w, "conv4", 0))
    up = lambda t, s: F.interpolate(t, scale_factor=s, mode="bilinear", align_corners=True)
    x1234 = torch.cat([x1, up(x2, 2), up(x3, 8), up(x4, 32)], dim=1)
    s = F.selu(_conv(x1234, w, "score_head.0", 0))
    s = F.selu(_conv(s, w, "score_head.2"))
    s = F.selu(_conv(s, w, "score_head.4"))
    score = torch.sigmoid(_conv(s, w, "score_head.6"))
    feat = F.normalize(x1234, p=2, dim=1)
    hh, ww = feat.shape[-2:]
    c = [pad[2], hh - pad[3], pad[0], ww - pad[1]]
    return feat[..., c[0]:c[1], c[2]:c[3]], score[..., c[0]:c[1], c[2]:c[3]]


def dkd(score_map: torch.Tensor, radius: int, scores_th: float, n_limit: int):
    """DKD.forward (:123-244), threshold mode with sub-pixel refinement. Returns (kpts in [-1,1], dispersity, score)."""
    _, _, h, w = score_map.shape
    nms = simple_nms(score_map, radius)
    nms[:, :, :radius, :] = 0
    nms[:, :, :, :radius] = 0
    nms[:, :, -radius:, :] = 0
    nms[:, :, :, -radius:] = 0
    if scores_th > 0:
        mask = nms > scores_th
        if mask.sum() == 0:
            mask = nms > score_map.reshape(1, -1).mean(dim=1).reshape(1, 1, 1, 1)
    else:
        mask = nms > score_map.reshape(1, -1).mean(dim=1).reshape(1, 1, 1, 1)
    scores_view = score_map.reshape(-1)
    idx = mask.reshape(-1).nonzero()[:, 0]
    if len(idx) > n_limit:
        sel = scores_view[idx].sort(descending=True)[1][:n_limit]
        idx = idx[sel]
    ks = 2 * radius + 1
    xs = torch.linspace(-radius, radius, ks)
    hw_grid = torch.stack(torch.meshgrid([xs, xs], indexing="ij")).view(2, -1).t()[:, [1, 0]]
    patches = F.unfold(score_map, kernel_size=ks, padding=radius)[0].t()  # (H*W, ks*ks)
    patch = patches[idx]
    xy_nms = torch.stack([idx % w, torch.div(idx, w, rounding_mode="trunc")], dim=1)
    max_v = patch.max(dim=1).values[:, None]
    x_exp = ((patch - max_v) / 0.1).exp()
    xy_res = x_exp @ hw_grid / x_exp.sum(dim=1)[:, None]
    d2 = torch.norm((hw_grid[None] - xy_res[:, None]) / radius, dim=-1) ** 2
    disp = (x_exp * d2).sum(dim=1) / x_exp.sum(dim=1)
    wh = torch.tensor([w - 1, h - 1])
    kxy = (xy_nms + xy_res) / wh * 2 - 1
    kscore = F.grid_sample(score_map, kxy.view(1, 1, -1, 2), mode="bilinear", align_corners=True)[0, 0, 0, :]
    return kxy, disp, kscore


def get_patches(tensor, corners, ps):
    """aliked.py:48-63."""
    c, h, w = tensor.shape
    corner = (corners - ps / 2 + 1).long()
    corner[:, 0] = corner[:, 0].clamp(min=0, max=w - 1 - ps)
    corner[:, 1] = corner[:, 1].clamp(min=0, max=h - 1 - ps)
    offset = torch.arange(0, ps)
    x, y = torch.meshgrid(offset, offset, indexing="ij")
    patches = torch.stack((x, y)).permute(2, 1, 0).unsqueeze(2)
    patches = patches.to(corner) + corner[None, None]
    pts = patches.reshape(-1, 2)
    sampled = tensor.permute(1, 2, 0)[tuple(pts.T)[::-1]]
    return sampled.reshape(ps, ps, -1, c).permute(2, 3, 0, 1)


def sddh(feat: torch.Tensor, kpts: torch.Tensor, w: dict, K: int, M: int):
    """SDDH.forward (:503-558) for one image: feat (1,C,H,W), kpts (N,2) in [-1,1] -> (N,C)."""
    _, c, h, wd = feat.shape
    wh = torch.tensor([[wd - 1, h - 1]])
    max_offset = max(h, wd) / 4.0
    kwh = (kpts / 2 + 0.5) * wh
    n = len(kpts)
    patch = get_patches(feat[0], kwh.long(), K)
    off = F.conv2d(patch, _t(w, "desc_head.offset_conv.0.weight"), _t(w, "desc_head.offset_conv.0.bias"))
    off = F.conv2d(F.selu(off), _t(w, "desc_head.offset_conv.2.weight"), _t(w, "desc_head.offset_conv.2.bias"))
    off = off.clamp(-max_offset, max_offset)[:, :, 0, 0].view(n, 2, M).permute(0, 2, 1)
    pos = kwh.unsqueeze(1) + off
    pos = (2.0 * pos / wh[None] - 1).reshape(1, n * M, 1, 2)
    f = F.grid_sample(feat, pos, mode="bilinear", align_corners=True)
    f = f.reshape(c, n, M, 1).permute(1, 0, 2, 3)
    f = torch.selu_(F.conv2d(f, _t(w, "desc_head.sf_conv.weight"))).squeeze(-1)
    d = torch.einsum("ncp,pcd->nd", f, _t(w, "desc_head.agg_weights"))
    return F.normalize(d, p=2.0, dim=1)


@torch.no_grad()
def extract(image: np.ndarray, w: dict, conf: dict | None = None, return_debug: bool = False) -> dict:
    """Oracle of ``AlikedExtractor._extract``: image float32 (H,W,3) RGB 0..255 (or (H,W) gray).
    Returns keypoints float32 (N,2) sub-pixel xy, descriptors float32 (128,N), scores float32 (N,) (= dispersity)."""
    conf = {**DEFAULT_CONF, **(conf or {})}
    cfg = CFGS[conf["model_name"]]
    if image.ndim == 2:
        x = torch.tensor(image[None][None] / 255.0, dtype=torch.float).repeat(1, 3, 1, 1)  # grayscale_to_rgb
    else:
        x = torch.tensor(image.transpose(2, 0, 1)[None] / 255.0, dtype=torch.float)
    feat, score = dense_maps(x, w, cfg)
    n_limit = conf["max_num_keypoints"] if conf["max_num_keypoints"] > 0 else N_LIMIT_MAX
    kxy, disp, kscore = dkd(score, conf["nms_radius"], conf["detection_threshold"], n_limit)
    desc = sddh(feat, kxy, w, cfg[5], cfg[6])
    _, _, h, wd = x.shape
    wh = torch.tensor([wd - 1, h - 1])
    out = {
        "keypoints": (wh * (kxy + 1) / 2.0).numpy().astype(np.float32),
        "descriptors": desc.t().contiguous().numpy().astype(np.float32),
        "scores": disp.numpy().astype(np.float32),  # sic: dispersity (quirk A.5)
    }
    if return_debug:
        out["_score_map"] = score[0, 0].numpy()
        out["_feature_map"] = feat[0].numpy()
        out["_kptscore"] = kscore.numpy()
    return out
