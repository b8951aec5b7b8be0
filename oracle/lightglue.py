"""LightGlue CPU oracle (test infrastructure only - see oracle/__init__.py).

Functional torch-fp32 restatement of the graph the reference runs for
``LightGlueMatcher._match_pairs``:

* adapter              : src/deep_image_matching/matchers/lightglue.py:8-66,102-125
* normalize_keypoints  : thirdparty/LightGlue/lightglue/lightglue.py:24-34
* positional encoding  : :57-70, rotary :41-54
* SelfBlock            : :129-159     CrossBlock : :162-211
* token confidence     : :73-83, stop test :593-604, thresholds :581-584
* point pruning        : :481-516, mask :586-591, min-kpts :606-610 / :318-323
* assignment           : :246-275     filter_matches : :281-297
* result assembly      : :540-579

Control flow follows the reference's CUDA+flash semantics (pruning only while
more than ``prune_min_kpts``=1536 points remain);
arithmetic is fp32 (``attn_half=True`` emulates the fp16 q/k/v cast of
lightglue.py:105-110 for the secondary "fast mode" comparison).

Weights: dict name -> np.float32 array using the reference's state_dict names
(SURVEY Appendix D).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_CONF = {
    "input_dim": 256,
    "descriptor_dim": 256,
    "n_layers": 9,
    "num_heads": 4,
    "depth_confidence": 0.95,
    "width_confidence": 0.99,
    "filter_threshold": 0.1,
    "prune_min_kpts": 1536,  # pruning_keypoint_thresholds["flash"], lightglue.py:318-323
    "attn_half": False,
}


def confidence_threshold(i: int, n_layers: int) -> float:
    """lightglue.py:581-584."""
    return float(np.clip(0.8 + 0.1 * np.exp(-4.0 * i / n_layers), 0, 1))


def seeded_weights(conf: dict | None = None, seed: int = 0, structured: bool = True) -> dict:
    """Deterministic LightGlue-architecture weights (single definition: dim_b200.weights.lightglue_seeded)."""
    from dim_b200.weights import lightglue_seeded

    c = {**DEFAULT_CONF, **(conf or {})}
    return lightglue_seeded(c["input_dim"], c["descriptor_dim"], c["n_layers"], c["num_heads"], seed, structured)


def _t(w, name):
    return torch.from_numpy(w[name])


def _linear(x, w, name):
    return F.linear(x, _t(w, name + ".weight"), _t(w, name + ".bias"))


def normalize_keypoints(kpts: torch.Tensor, size: torch.Tensor) -> torch.Tensor:
    """lightglue.py:24-34. ``size`` is whatever the caller passed as image_size ([H,W] in DIM)."""
    size = size.to(kpts)
    shift = size / 2
    scale = size.max(-1).values / 2
    return (kpts - shift[None, :]) / scale


def posenc(kn: torch.Tensor, w: dict) -> torch.Tensor:
    """(N,2) -> (2,N,hd): [cos,sin] with each frequency repeated twice (:57-70)."""
    proj = F.linear(kn, _t(w, "posenc.Wr.weight"))
    emb = torch.stack([torch.cos(proj), torch.sin(proj)], 0)
    return emb.repeat_interleave(2, dim=-1)


def _rotate_half(x):
    x = x.unflatten(-1, (-1, 2))
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).flatten(start_dim=-2)


def _rotary(enc, t):
    """enc (2,N,hd), t (h,N,hd)."""
    return t * enc[0][None] + _rotate_half(t) * enc[1][None]


def _attention(q, k, v, attn_half: bool):
    """(h,N,hd),(h,M,hd),(h,M,hd) -> (h,N,hd); scale hd^-0.5 (:102-126)."""
    if q.shape[-2] == 0 or k.shape[-2] == 0:
        return q.new_zeros((*q.shape[:-1], v.shape[-1]))
    if attn_half:
        q, k, v = q.half().float(), k.half().float(), v.half().float()
    s = torch.einsum("hid,hjd->hij", q, k) * (q.shape[-1] ** -0.5)
    return torch.einsum("hij,hjd->hid", F.softmax(s, -1), v)


def _ffn(x, msg, w, p):
    y = _linear(torch.cat([x, msg], -1), w, p + ".ffn.0")
    y = F.layer_norm(y, (y.shape[-1],), _t(w, p + ".ffn.1.weight"), _t(w, p + ".ffn.1.bias"), 1e-5)
    y = F.gelu(y)
    return x + _linear(y, w, p + ".ffn.3")


def self_block(x, enc, w, p, h, attn_half):
    """(N,d) -> (N,d). lightglue.py:146-159 (Wqkv output interleaved as (h, hd, 3))."""
    n, d = x.shape
    qkv = _linear(x, w, p + ".Wqkv").unflatten(-1, (h, -1, 3)).transpose(0, 1)  # (h,N,hd,3)
    q, k, v = qkv[..., 0], qkv[..., 1], qkv[..., 2]
    q, k = _rotary(enc, q), _rotary(enc, k)
    ctx = _attention(q, k, v, attn_half)  # (h,N,hd)
    msg = _linear(ctx.transpose(0, 1).flatten(start_dim=-2), w, p + ".out_proj")
    return _ffn(x, msg, w, p)


def cross_block(x0, x1, w, p, h, attn_half):
    """lightglue.py:186-211 (flash branch: m0 = attn(qk0, qk1, v1), m1 = attn(qk1, qk0, v0))."""
    def heads(t):
        return t.unflatten(-1, (h, -1)).transpose(0, 1)

    qk0, qk1 = heads(_linear(x0, w, p + ".to_qk")), heads(_linear(x1, w, p + ".to_qk"))
    v0, v1 = heads(_linear(x0, w, p + ".to_v")), heads(_linear(x1, w, p + ".to_v"))
    m0 = _attention(qk0, qk1, v1, attn_half)
    m1 = _attention(qk1, qk0, v0, attn_half)
    m0 = _linear(m0.transpose(0, 1).flatten(start_dim=-2), w, p + ".to_out")
    m1 = _linear(m1.transpose(0, 1).flatten(start_dim=-2), w, p + ".to_out")
    return _ffn(x0, m0, w, p), _ffn(x1, m1, w, p)


def log_assignment(d0, d1, w, i):
    """(M,d),(N,d) -> (M+1,N+1) log assignment. lightglue.py:246-275."""
    p = f"log_assignment.{i}"
    dim = d0.shape[-1]
    md0 = _linear(d0, w, p + ".final_proj") / dim**0.25
    md1 = _linear(d1, w, p + ".final_proj") / dim**0.25
    sim = md0 @ md1.t()
    z0 = _linear(d0, w, p + ".matchability")
    z1 = _linear(d1, w, p + ".matchability")
    m, n = sim.shape
    cert = F.logsigmoid(z0) + F.logsigmoid(z1).t()
    s0 = F.log_softmax(sim, 1)
    s1 = F.log_softmax(sim.t().contiguous(), 1).t()
    scores = sim.new_zeros((m + 1, n + 1))
    scores[:m, :n] = s0 + s1 + cert
    scores[:-1, -1] = F.logsigmoid(-z0.squeeze(-1))
    scores[-1, :-1] = F.logsigmoid(-z1.squeeze(-1))
    return scores


def filter_matches(scores: torch.Tensor, th: float):
    """lightglue.py:281-297 (unbatched)."""
    inner = scores[:-1, :-1]
    max0, max1 = inner.max(1), inner.max(0)
    m0, m1 = max0.indices, max1.indices
    i0 = torch.arange(m0.shape[0])
    i1 = torch.arange(m1.shape[0])
    mutual0 = i0 == m1.gather(0, m0)
    mutual1 = i1 == m0.gather(0, m1)
    max0_exp = max0.values.exp()
    zero = max0_exp.new_tensor(0)
    ms0 = torch.where(mutual0, max0_exp, zero)
    ms1 = torch.where(mutual1, ms0.gather(0, m1), zero)
    valid0 = mutual0 & (ms0 > th)
    valid1 = mutual1 & valid0.gather(0, m1)
    m0 = torch.where(valid0, m0, -1)
    m1 = torch.where(valid1, m1, -1)
    return m0, m1, ms0, ms1


def features_to_lg(feats: dict):
    """featuresDict2Lightglue (matchers/lightglue.py:8-66): (D,N)->(N,D) decided by N, float32."""
    kpts = np.asarray(feats["keypoints"])
    desc = np.asarray(feats["descriptors"])
    if kpts.ndim != 2 or kpts.shape[1] != 2:
        raise ValueError(f"Invalid keypoints shape: {kpts.shape}")
    n = kpts.shape[0]
    if desc.ndim != 2:
        raise ValueError(f"Invalid descriptors shape: {desc.shape}")
    if desc.shape[1] == n and desc.shape[0] != n:
        desc = desc.T
    elif desc.shape[0] == n:
        pass
    else:
        raise ValueError(f"Descriptor / keypoint mismatch: descriptors={desc.shape}, keypoints={kpts.shape}")
    size = feats.get("image_size")
    return (
        torch.as_tensor(np.ascontiguousarray(kpts), dtype=torch.float32),
        torch.as_tensor(np.ascontiguousarray(desc), dtype=torch.float32),
        None if size is None else torch.as_tensor(np.asarray(size), dtype=torch.float32),
    )


@torch.no_grad()
def match(feats0: dict, feats1: dict, w: dict, conf: dict | None = None, return_debug: bool = False) -> dict:
    """Oracle of LightGlue._forward on one pair (batch 1).

    Returns dict(matches int64 (S,2), scores float32 (S,), stop int, prune0, prune1,
    matches0, matching_scores0 ...).
    """
    c = {**DEFAULT_CONF, **(conf or {})}
    L, h = c["n_layers"], c["num_heads"]
    k0, d0, s0 = features_to_lg(feats0)
    k1, d1, s1 = features_to_lg(feats1)
    m, n = k0.shape[0], k1.shape[0]
    assert d0.shape[-1] == c["input_dim"] and d1.shape[-1] == c["input_dim"]
    if s0 is None:  # lightglue.py:26-27
        s0 = 1 + k0.max(-2).values - k0.min(-2).values
    if s1 is None:
        s1 = 1 + k1.max(-2).values - k1.min(-2).values
    kn0, kn1 = normalize_keypoints(k0, s0), normalize_keypoints(k1, s1)
    if c["input_dim"] != c["descriptor_dim"]:
        d0, d1 = _linear(d0, w, "input_proj"), _linear(d1, w, "input_proj")
    e0, e1 = posenc(kn0, w), posenc(kn1, w)

    do_stop = c["depth_confidence"] > 0
    do_prune = c["width_confidence"] > 0
    ind0, ind1 = torch.arange(m), torch.arange(n)
    prune0, prune1 = torch.ones(m, dtype=torch.long), torch.ones(n, dtype=torch.long)
    thr = torch.tensor([confidence_threshold(i, L) for i in range(L)], dtype=torch.float32)
    tok0 = tok1 = None
    dbg = {"desc0": [], "desc1": [], "n0": [], "n1": []}
    i = 0
    for i in range(L):
        if d0.shape[0] == 0 or d1.shape[0] == 0:
            break
        p = f"transformers.{i}."
        d0 = self_block(d0, e0, w, p + "self_attn", h, c["attn_half"])
        d1 = self_block(d1, e1, w, p + "self_attn", h, c["attn_half"])
        d0, d1 = cross_block(d0, d1, w, p + "cross_attn", h, c["attn_half"])
        if return_debug:
            dbg["desc0"].append(d0.numpy().copy())
            dbg["desc1"].append(d1.numpy().copy())
            dbg["n0"].append(d0.shape[0])
            dbg["n1"].append(d1.shape[0])
        if i == L - 1:
            continue
        if do_stop:
            tp = f"token_confidence.{i}.token.0"
            tok0 = torch.sigmoid(_linear(d0, w, tp)).squeeze(-1)
            tok1 = torch.sigmoid(_linear(d1, w, tp)).squeeze(-1)
            conf_all = torch.cat([tok0, tok1], -1)
            ratio = 1.0 - (conf_all < thr[i]).float().sum() / (m + n)
            if ratio > c["depth_confidence"]:
                break
        if do_prune and d0.shape[0] > c["prune_min_kpts"]:
            sc = torch.sigmoid(_linear(d0, w, f"log_assignment.{i}.matchability")).squeeze(-1)
            keep = sc > (1 - c["width_confidence"])
            if tok0 is not None:
                keep |= tok0 <= thr[i]
            kidx = torch.where(keep)[0]
            ind0, d0, e0 = ind0[kidx], d0[kidx], e0[:, kidx]
            prune0[ind0] += 1
        if do_prune and d1.shape[0] > c["prune_min_kpts"]:
            sc = torch.sigmoid(_linear(d1, w, f"log_assignment.{i}.matchability")).squeeze(-1)
            keep = sc > (1 - c["width_confidence"])
            if tok1 is not None:
                keep |= tok1 <= thr[i]
            kidx = torch.where(keep)[0]
            ind1, d1, e1 = ind1[kidx], d1[kidx], e1[:, kidx]
            prune1[ind1] += 1

    if d0.shape[0] == 0 or d1.shape[0] == 0:
        return {
            "matches": np.zeros((0, 2), np.int64),
            "scores": np.zeros((0,), np.float32),
            "stop": i + 1,
            "prune0": prune0.numpy(),
            "prune1": prune1.numpy(),
        }
    la = log_assignment(d0, d1, w, i)
    m0, m1, ms0, ms1 = filter_matches(la, c["filter_threshold"])
    valid = m0 > -1
    a = torch.where(valid)[0]
    b = m0[valid]
    matches = torch.stack([ind0[a], ind1[b]], -1)
    out = {
        "matches": matches.numpy().astype(np.int64),
        "scores": ms0[valid].numpy().astype(np.float32),
        "stop": i + 1,
        "prune0": prune0.numpy() if do_prune else np.full(m, L),
        "prune1": prune1.numpy() if do_prune else np.full(n, L),
        "n_final0": int(d0.shape[0]),
        "n_final1": int(d1.shape[0]),
    }
    if return_debug:
        out["_dbg"] = dbg
        out["_log_assignment"] = la.numpy()
        out["_ind0"], out["_ind1"] = ind0.numpy(), ind1.numpy()
    return out
