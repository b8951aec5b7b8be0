"""SuperGlue CPU oracle (test infrastructure only - see oracle/__init__.py).  No CUDA path exists for it yet: this is the
pinned specification for the SURVEY 8(f) "next" row, whose attention shape (256 / 4 heads x 64) is the one the tensor-core
kernels of csrc/lightglue.cu are built for.

Functional torch-fp32 restatement of what ``SuperGlueMatcher._match_pairs`` runs:

* adapter : src/deep_image_matching/matchers/superglue.py:8-106 (``features_2_sg`` builds ``image{i}`` as an empty
            (1,1,H,W) tensor from ``image_size`` = [H,W]; FeaturesDict descriptors are already (D,N); ``scores`` are required)
* model   : thirdparty/SuperGluePretrainedNetwork/models/superglue.py:51-305 (normalize_keypoints :63-70, KeypointEncoder
            :73-84, attention :87-93, MultiHeadedAttention :96-117 with the (dim, heads) channel interleave of ``view``,
            AttentionalPropagation :120-129, AttentionalGNN :132-152, log-space Sinkhorn :155-187, matching :278-296)

Reproduced quirk: the plugin builds its config from ``self._default_conf`` (empty in MatcherBase), not from its own
``default_config`` (superglue.py:55-60,72), so the model's defaults apply: 100 Sinkhorn iterations, match_threshold 0.2.

Weights: dict name -> np.float32 array with the reference's state_dict names.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_CONF = {  # superglue.py:213-220
    "descriptor_dim": 256,
    "keypoint_encoder": [32, 64, 128, 256],
    "GNN_layers": ["self", "cross"] * 9,
    "sinkhorn_iterations": 100,
    "match_threshold": 0.2,
    "num_heads": 4,
}


def _t(w, k):
    return torch.from_numpy(np.ascontiguousarray(w[k]))


def _conv1(x, w, p):
    """Conv1d(kernel 1) on (C,N)."""
    return _t(w, p + ".weight")[:, :, 0] @ x + _t(w, p + ".bias")[:, None]


def _bn(x, w, p):
    return F.batch_norm(x[None], _t(w, p + ".running_mean"), _t(w, p + ".running_var"), _t(w, p + ".weight"), _t(w, p + ".bias"),
                        False, 0.0, 1e-5)[0]


def _mlp(x, w, p, n_layers):
    """MLP (:51-60): conv1d, then BN + ReLU after every layer but the last; Sequential indices 0,1,2 | 3,4,5 | ..."""
    for i in range(n_layers):
        x = _conv1(x, w, f"{p}.{3 * i}")
        if i < n_layers - 1:
            x = F.relu(_bn(x, w, f"{p}.{3 * i + 1}"))
    return x


def normalize_keypoints(kpts, h, w):
    size = torch.tensor([[float(w), float(h)]])
    center = size / 2
    scaling = size.max(1, keepdim=True).values * 0.7
    return (kpts - center) / scaling


def _attention_block(x, src, w, p, heads):
    """MultiHeadedAttention (:96-117) on (D,N) / (D,M): channel c = d * heads + h."""
    d_model = x.shape[0]
    dim = d_model // heads
    q = _conv1(x, w, p + ".proj.0").view(dim, heads, -1)
    k = _conv1(src, w, p + ".proj.1").view(dim, heads, -1)
    v = _conv1(src, w, p + ".proj.2").view(dim, heads, -1)
    scores = torch.einsum("dhn,dhm->hnm", q, k) / dim**0.5
    prob = F.softmax(scores, dim=-1)
    out = torch.einsum("hnm,dhm->dhn", prob, v)
    return _conv1(out.contiguous().view(d_model, -1), w, p + ".merge")


def _propagate(x, src, w, p, heads):
    msg = _attention_block(x, src, w, p + ".attn", heads)
    return _mlp(torch.cat([x, msg], 0), w, p + ".mlp", 2)


def log_optimal_transport(scores, alpha, iters):
    """(:169-187) for one pair: scores (M,N) -> (M+1,N+1)."""
    m, n = scores.shape
    ms, ns = torch.tensor(float(m)), torch.tensor(float(n))
    couplings = torch.cat([torch.cat([scores, alpha.expand(m, 1)], -1), torch.cat([alpha.expand(1, n), alpha.expand(1, 1)], -1)], 0)
    norm = -(ms + ns).log()
    log_mu = torch.cat([norm.expand(m), ns.log()[None] + norm])
    log_nu = torch.cat([norm.expand(n), ms.log()[None] + norm])
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(couplings + v[None, :], dim=1)
        v = log_nu - torch.logsumexp(couplings + u[:, None], dim=0)
    return couplings + u[:, None] + v[None, :] - norm


@torch.no_grad()
def match(feats0: dict, feats1: dict, w: dict, conf: dict | None = None) -> dict:
    """Oracle of SuperGlue.forward driven as SuperGlueMatcher does.  feats: keypoints (N,2), descriptors (256,N), scores (N,),
    image_size [H,W].  Returns matches int64 (S,2) (correspondence_matrix_from_matches0), matches0, matching_scores0."""
    c = {**DEFAULT_CONF, **(conf or {})}
    k0, k1 = torch.from_numpy(np.asarray(feats0["keypoints"], np.float32)), torch.from_numpy(np.asarray(feats1["keypoints"], np.float32))
    d0, d1 = torch.from_numpy(np.asarray(feats0["descriptors"], np.float32)), torch.from_numpy(np.asarray(feats1["descriptors"], np.float32))
    s0, s1 = torch.from_numpy(np.asarray(feats0["scores"], np.float32)), torch.from_numpy(np.asarray(feats1["scores"], np.float32))
    m, n = k0.shape[0], k1.shape[0]
    if m == 0 or n == 0:
        return {"matches": np.zeros((0, 2), np.int64), "matches0": np.full(m, -1, np.int64), "matching_scores0": np.zeros(m, np.float32)}
    (h0, w0), (h1, w1) = [int(v) for v in feats0["image_size"]], [int(v) for v in feats1["image_size"]]
    kn0, kn1 = normalize_keypoints(k0, h0, w0), normalize_keypoints(k1, h1, w1)
    nk = len(c["keypoint_encoder"]) + 1
    d0 = d0 + _mlp(torch.cat([kn0.t(), s0[None]], 0), w, "kenc.encoder", nk)
    d1 = d1 + _mlp(torch.cat([kn1.t(), s1[None]], 0), w, "kenc.encoder", nk)
    for i, name in enumerate(c["GNN_layers"]):
        src0, src1 = (d1, d0) if name == "cross" else (d0, d1)
        p = f"gnn.layers.{i}"
        delta0, delta1 = _propagate(d0, src0, w, p, c["num_heads"]), _propagate(d1, src1, w, p, c["num_heads"])
        d0, d1 = d0 + delta0, d1 + delta1
    md0, md1 = _conv1(d0, w, "final_proj"), _conv1(d1, w, "final_proj")
    scores = (md0.t() @ md1) / c["descriptor_dim"] ** 0.5
    Z = log_optimal_transport(scores, _t(w, "bin_score").reshape(()), c["sinkhorn_iterations"])
    inner = Z[:-1, :-1]
    max0, max1 = inner.max(1), inner.max(0)
    i0, i1 = max0.indices, max1.indices
    mutual0 = torch.arange(m) == i1.gather(0, i0)
    ms0 = torch.where(mutual0, max0.values.exp(), torch.tensor(0.0))
    valid0 = mutual0 & (ms0 > c["match_threshold"])
    matches0 = torch.where(valid0, i0, torch.tensor(-1))
    a = torch.where(valid0)[0]
    return {"matches": torch.stack([a, matches0[a]], -1).numpy().astype(np.int64), "matches0": matches0.numpy().astype(np.int64),
            "matching_scores0": ms0.numpy().astype(np.float32)}


def seeded_weights(seed: int = 0, n_gnn: int = 18) -> dict:
    """Deterministic SuperGlue-architecture weights (numpy PCG64), structured so that matching fires: small residual updates,
    final_proj = 12 * (I + 0.05 W) (scores ~ 9 * cosine), non-trivial BatchNorm statistics."""
    rng = np.random.default_rng(seed)
    w = {}

    def conv(name, cout, cin, scale=1.0, bias=0.1):
        w[name + ".weight"] = (scale * rng.standard_normal((cout, cin, 1)) / np.sqrt(cin)).astype(np.float32)
        w[name + ".bias"] = (bias * rng.standard_normal(cout)).astype(np.float32)

    def bn(name, c):
        w[name + ".weight"] = (1 + 0.1 * rng.standard_normal(c)).astype(np.float32)
        w[name + ".bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
        w[name + ".running_mean"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
        w[name + ".running_var"] = (1 + 0.2 * rng.random(c)).astype(np.float32)

    ch = [3, 32, 64, 128, 256, 256]
    for i in range(5):
        conv(f"kenc.encoder.{3 * i}", ch[i + 1], ch[i], scale=0.3 if i == 4 else 1.0)
        if i < 4:
            bn(f"kenc.encoder.{3 * i + 1}", ch[i + 1])
    for i in range(n_gnn):
        p = f"gnn.layers.{i}"
        conv(p + ".attn.merge", 256, 256)
        for j in range(3):
            conv(p + f".attn.proj.{j}", 256, 256, scale=2.0 if j < 2 else 1.0)
        conv(p + ".mlp.0", 512, 512)
        bn(p + ".mlp.1", 512)
        conv(p + ".mlp.3", 256, 512, scale=0.15)
    fw = 12.0 * (np.eye(256) + 0.05 * rng.standard_normal((256, 256)) / 16.0)
    w["final_proj.weight"] = fw[:, :, None].astype(np.float32)
    w["final_proj.bias"] = (0.05 * rng.standard_normal(256)).astype(np.float32)
    w["bin_score"] = np.array(1.0, np.float32)
    return w
