"""Weight loading for the plugins (numpy dicts keyed by the reference's state_dict names)."""
from __future__ import annotations

import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)
DATA = os.path.join(_HERE, "data")  # checkpoints the reference vendors, converted to .npz (oracle/gen_golden.py)


def load_npz(path) -> dict:
    z = np.load(path)
    return {k: z[k].astype(np.float32) for k in z.files}


def superpoint_v1() -> dict:
    """superpoint_v1 weights (converted from the reference's vendored superpoint_v1.pth by oracle/gen_golden.py)."""
    for p in (os.environ.get("DIMB_SUPERPOINT_WEIGHTS"), os.path.join(DATA, "superpoint_v1_weights.npz")):
        if p and os.path.exists(p):
            return load_npz(p)
    raise FileNotFoundError("superpoint_v1 weights not found (set DIMB_SUPERPOINT_WEIGHTS)")


def aliked(model_name: str = "aliked-n16rot") -> dict:
    """ALIKED weights by model name (the reference downloads ``<model_name>.pth`` from the ALIKED release, aliked.py:581-587;
    here the vendored thirdparty/ALIKED/models/aliked-n16.pth / aliked-n16rot.pth converted by oracle/gen_golden.py).
    Lookup: DIMB_ALIKED_WEIGHTS_<MODEL> (e.g. DIMB_ALIKED_WEIGHTS_N16ROT), DIMB_ALIKED_WEIGHTS, then the packaged file."""
    tag = model_name.replace("aliked-", "")
    for p in (os.environ.get("DIMB_ALIKED_WEIGHTS_" + tag.upper()), os.environ.get("DIMB_ALIKED_WEIGHTS"),
              os.path.join(DATA, f"aliked_{tag}_weights.npz")):
        if p and os.path.exists(p):
            return load_npz(p) if p.endswith(".npz") else from_torch_checkpoint(p)
    raise FileNotFoundError(f"{model_name} weights not found (set DIMB_ALIKED_WEIGHTS_{tag.upper()})")


def aliked_n16rot() -> dict:
    return aliked("aliked-n16rot")


def from_torch_checkpoint(path) -> dict:
    import torch

    sd = torch.load(str(path), map_location="cpu")
    return {k: v.numpy().astype(np.float32) for k, v in sd.items() if v.dtype.is_floating_point}


def lightglue_seeded(input_dim: int = 256, descriptor_dim: int = 256, n_layers: int = 9, num_heads: int = 4,
                     seed: int = 0, structured: bool = True) -> dict:
    """Deterministic, platform-independent LightGlue-architecture weights.

    No pretrained LightGlue checkpoint exists offline, so parity is
    architecture-level: numpy PCG64 draws with PyTorch-like fan-in scaling.
    Plain random weights give a uniform assignment (0 matches), so by default
    the draw is *structured* to behave like a trained network: modest residual
    updates, final_proj ~ 13*(I + noise) so true correspondences win the double
    softmax, matchability / token-confidence heads with enough spread and a
    per-layer bias ramp that the early-exit and point-pruning branches fire.
    The same generator feeds the reference model (oracle/gen_golden.py), the
    oracle and the CUDA path (bench.py uses it as the random-init LightGlue).
    """
    d, din, L = descriptor_dim, input_dim, n_layers
    hd = d // num_heads
    rng = np.random.Generator(np.random.PCG64(seed))
    w = {}

    def lin(name, out_f, in_f):
        b = 1.0 / math.sqrt(in_f)
        w[name + ".weight"] = rng.uniform(-b, b, (out_f, in_f)).astype(np.float32)
        w[name + ".bias"] = rng.uniform(-b, b, (out_f,)).astype(np.float32)

    w["posenc.Wr.weight"] = rng.standard_normal((hd // 2, 2)).astype(np.float32)
    if din != d:
        lin("input_proj", d, din)
    for i in range(L):
        p = f"transformers.{i}."
        lin(p + "self_attn.Wqkv", 3 * d, d)
        lin(p + "self_attn.out_proj", d, d)
        for blk in ("self_attn", "cross_attn"):
            lin(p + blk + ".ffn.0", 2 * d, 2 * d)
            w[p + blk + ".ffn.1.weight"] = (1.0 + 0.1 * rng.standard_normal(2 * d)).astype(np.float32)
            w[p + blk + ".ffn.1.bias"] = (0.1 * rng.standard_normal(2 * d)).astype(np.float32)
            lin(p + blk + ".ffn.3", d, 2 * d)
        lin(p + "cross_attn.to_qk", d, d)
        lin(p + "cross_attn.to_v", d, d)
        lin(p + "cross_attn.to_out", d, d)
        lin(f"log_assignment.{i}.matchability", 1, d)
        lin(f"log_assignment.{i}.final_proj", d, d)
        if i < L - 1:
            lin(f"token_confidence.{i}.token.0", 1, d)
    if structured:
        if din != d:  # keep projected descriptors near unit norm and similarity-preserving
            w["input_proj.weight"] = (w["input_proj.weight"] * math.sqrt(3.0 * din / d) * 1.0).astype(np.float32)
        for i in range(L):
            for blk in ("self_attn", "cross_attn"):
                p = f"transformers.{i}.{blk}.ffn.3."
                w[p + "weight"] *= np.float32(0.15)
                w[p + "bias"] *= np.float32(0.15)
            p = f"log_assignment.{i}.final_proj."
            w[p + "weight"] = (13.0 * (np.eye(d, dtype=np.float32) + 0.3 * w[p + "weight"])).astype(np.float32)
            w[p + "bias"] *= np.float32(13.0)
            p = f"log_assignment.{i}.matchability."
            w[p + "weight"] *= np.float32(40.0)
            w[p + "bias"][:] = 4.0
            if i < L - 1:
                p = f"token_confidence.{i}.token.0."
                w[p + "weight"] *= np.float32(12.0)
                w[p + "bias"][:] = -3.0 + 1.2 * i
    return w


