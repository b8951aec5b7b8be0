"""Hardware probe: tcgen05.mma with the A operand in tensor memory (the P matrix of the attention kernel lg_attn5_kernel).
Thread r stores row r of A as packed half2 words (word c = elements 2c, 2c+1) with tcgen05.st.32x32b; four 16-deep MMAs read it at
column offsets 0 / 8 / 16 / 24.  With B = identity the result IS the A tile as the tensor core saw it.  Writes
gpurun_out/probe_tmem_a.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from dim_b200 import _native
    ctx = _native.SelfTest(0)
    out = {}
    # two fp16-exact images of the A tile: one names the row, one names the k index (B = identity: the result IS A as the MMA saw it)
    C = np.zeros((128, 64), np.float32)
    eye = np.eye(64, dtype=np.float32)
    for name, A in (("rows", np.broadcast_to(np.arange(128, dtype=np.float32)[:, None], (128, 64)).copy()),
                    ("k", np.broadcast_to(np.arange(64, dtype=np.float32)[None, :], (128, 64)).copy())):
        rc = ctx.lib.dimb_probe_tmem_a(ctx.h, _native._ptr(A), _native._ptr(eye), _native._ptr(C))
        out["identity_" + name] = {"rc": rc, "matches": bool(rc == 0 and np.array_equal(C, A))}
    out["identity"] = {"matches": out["identity_rows"]["matches"] and out["identity_k"]["matches"]}
    rng = np.random.default_rng(0)
    A2 = rng.standard_normal((128, 64)).astype(np.float16).astype(np.float32)
    B2 = rng.standard_normal((64, 64)).astype(np.float16).astype(np.float32)
    rc = ctx.lib.dimb_probe_tmem_a(ctx.h, _native._ptr(A2), _native._ptr(B2), _native._ptr(C))
    err = float(np.abs(C - A2 @ B2.T).max()) if rc == 0 else None
    out["random"] = {"rc": rc, "max_abs_err_vs_fp32": err, "matches": bool(rc == 0 and err < 1e-3)}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_tmem_a.json"), "w"), indent=1)
    return 0 if out["identity"]["matches"] and out["random"]["matches"] else 1


if __name__ == "__main__":
    sys.exit(main())
