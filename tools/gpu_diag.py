"""GPU bring-up diagnostics (run under gpurun): each stage is a separate process with a timeout so that a
trapped kernel cannot take the other stages down.   python tools/gpu_diag.py [stage ...]"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STAGES = ["gemm_simt", "gemm_tc", "sp_simt", "sp_tc", "lg_simt", "lg_tc", "nn", "time"]  # extra: sp_batch


def stage_gemm(tc: bool):
    import numpy as np
    from dim_b200 import _native
    ctx = _native.Context(0, tensor_path=tc)
    rng = np.random.default_rng(0)
    for prec in ("exact", "fast"):
        ctx.set_precision(prec)
        for (M, N, K, bn) in [(128, 128, 64, 128), (256, 256, 256, 128), (300, 200, 128, 64), (128, 256, 576, 256), (1000, 768, 512, 128)]:
            A = rng.standard_normal((M, K)).astype(np.float32)
            B = rng.standard_normal((N, K)).astype(np.float32)
            C = ctx.selftest_gemm(A, B, bn)
            ref = A.astype(np.float64) @ B.astype(np.float64).T
            err = np.abs(C - ref).max() / np.abs(ref).max()
            print(f"  gemm tc={tc} {prec} M{M} N{N} K{K} bn{bn}: rel err {err:.3e}", flush=True)


def stage_sp(tc: bool):
    import numpy as np
    from dim_b200 import _native, synthetic, weights
    from oracle import superpoint as o_sp
    ctx = _native.Context(0, tensor_path=tc)
    w = weights.superpoint_v1()
    conf = {"nms_radius": 3, "keypoint_threshold": 0.0005, "max_keypoints": 512}
    size = 256
    g0, g1 = synthetic.synthetic_pair(1, size)
    for prec in ("exact", "fast"):
        ctx.set_precision(prec)
        sp = _native.SuperPointNet(ctx, w, max_batch=2, max_height=size, max_width=size, **conf)
        t = time.time()
        feats = sp.extract(np.stack([g0, g1]))
        dt = time.time() - t
        ref = o_sp.extract(g0, w, conf, return_debug=True)
        h = size // 8
        feat = sp.debug_read(2, (h, h, 128))
        rf = ref["_feat"][0].transpose(1, 2, 0)
        print(f"  sp tc={tc} {prec}: encoder max|d|={np.abs(feat - rf).max():.3e} (ref max {np.abs(rf).max():.2f}) t={dt:.2f}s", flush=True)
        sc = sp.debug_read(0, (size, size))
        print(f"    dense scores max|d|={np.abs(sc - ref['_dense_scores']).max():.3e}")
        nms = sp.debug_read(1, (size, size))
        print(f"    nms map mismatching pixels={(np.abs(nms - ref['_nms']) > 1e-4).sum()} nonzero ref={(ref['_nms'] > 0).sum()} ours={(nms > 0).sum()}")
        dd = sp.debug_read(3, (h, h, 256))
        rd = ref["_dense_desc"][0].transpose(1, 2, 0)
        ddn = dd / np.maximum(np.linalg.norm(dd, axis=2, keepdims=True), 1e-12)
        print(f"    dense desc (normalised) max|d|={np.abs(ddn - rd).max():.3e}")
        f = feats[0]
        a, b = o_sp.canonical_order(f), o_sp.canonical_order(ref)
        same_n = len(a) == len(b)
        ka = {tuple(k) for k in f["keypoints"].astype(int)}
        kb = {tuple(k) for k in ref["keypoints"].astype(int)}
        print(f"    keypoints ours={len(a)} ref={len(b)} common={len(ka & kb)}")
        if same_n and ka == kb:
            print(f"    scores max|d|={np.abs(f['scores'][a] - ref['scores'][b]).max():.3e} desc max|d|={np.abs(f['descriptors'][:, a] - ref['descriptors'][:, b]).max():.3e}", flush=True)


def stage_sp_batch():
    """Every image of a batch, with and without the top-k branch, against the oracle (set differences + margins)."""
    import numpy as np
    from dim_b200 import _native, synthetic, weights
    from oracle import superpoint as o_sp
    ctx = _native.Context(0)
    w = weights.superpoint_v1()
    for size, mk in [(256, 512), (256, 200), (256, -1), (512, 1024)]:
        conf = {"nms_radius": 3, "keypoint_threshold": 0.0005, "max_keypoints": mk}
        g = np.stack(list(synthetic.synthetic_pair(1, size)) + list(synthetic.synthetic_pair(2, size)))
        sp = _native.SuperPointNet(ctx, w, max_batch=4, max_height=size, max_width=size, **conf)
        feats = sp.extract(g)
        for b in range(4):
            ref = o_sp.extract(g[b], w, conf, return_debug=True)
            f = feats[b]
            ko = {tuple(k) for k in f["keypoints"].astype(int)}
            kr = {tuple(k) for k in ref["keypoints"].astype(int)}
            cand = int((ref["_nms"] > conf["keypoint_threshold"]).sum())
            cut = float(ref["scores"].min()) if len(ref["scores"]) else 0
            d = sorted(ko ^ kr)
            margins = [(k, float(ref["_nms"][k[1], k[0]]) - cut) for k in d[:6]]
            print(f"  size {size} mk {mk} img {b}: ours {len(ko)} ref {len(kr)} cand~{cand} diff {len(d)} cut {cut:.6f} margins {margins}", flush=True)


def stage_lg(tc: bool):
    import numpy as np
    from dim_b200 import _native, weights
    from oracle import lightglue as o_lg
    from oracle.gen_golden import lg_pair
    ctx = _native.Context(0, tensor_path=tc)
    for prec in ("exact", "fast"):
        ctx.set_precision(prec)
        for (name, seed, m, n, over) in [("fixed", 1, 300, 260, {"depth_confidence": -1, "width_confidence": -1}), ("adaptive", 2, 512, 400, {}),
                                         ("prune", 7, 1700, 1650, {"depth_confidence": -1})]:
            conf = {**o_lg.DEFAULT_CONF, **over}
            w = weights.lightglue_seeded(seed=seed)
            f0, f1 = lg_pair(seed, m, n, 256, (768, 1024))
            lg = _native.LightGlueNet(ctx, w, depth_confidence=conf["depth_confidence"], width_confidence=conf["width_confidence"],
                                      max_pairs=1, max_kpts=max(m, n))
            t = time.time()
            out = lg.match([({**f0, "_layout": 0}, {**f1, "_layout": 0})])[0]
            dt = time.time() - t
            exp = o_lg.match(f0, f1, w, conf, return_debug=True)
            sa = {tuple(x) for x in out["matches"]}
            sb = {tuple(x) for x in exp["matches"]}
            msg = f"  lg tc={tc} {prec} {name}: stop {out['stop']}/{exp['stop']} matches {len(sa)}/{len(sb)} common {len(sa & sb)} t={dt:.2f}s"
            if sa == sb and len(sa):
                msg += f" score max|d|={np.abs(out['scores'] - exp['scores']).max():.3e}"
            NP = lg.NP
            xf = lg.debug_read(0, 0, (NP, 256))
            if out["stop"] == exp["stop"] and exp.get("n_final0") == exp["_dbg"]["n0"][-1]:
                ref_x = exp["_dbg"]["desc0"][-1]
                msg += f" final desc0 max|d|={np.abs(xf[:ref_x.shape[0]] - ref_x).max():.3e} (max {np.abs(ref_x).max():.2f})"
            print(msg, flush=True)


def stage_nn():
    import numpy as np
    from dim_b200 import _native
    from oracle import nn_match as o_nn
    rng = np.random.default_rng(5)
    for tc in (False, True):
        ctx = _native.Context(0, tensor_path=tc)
        for (n0, n1) in [(700, 650), (512, 777)]:
            a = rng.standard_normal((128, n0)).astype(np.float32); a /= np.linalg.norm(a, axis=0)
            b = rng.standard_normal((128, n1)).astype(np.float32); b /= np.linalg.norm(b, axis=0)
            k = min(n0, n1) // 2
            b[:, :k] = a[:, rng.permutation(n0)[:k]] + 0.3 * rng.standard_normal((128, k)).astype(np.float32)
            b /= np.linalg.norm(b, axis=0)
            a = a.astype(np.float16).astype(np.float32); b = b.astype(np.float16).astype(np.float32)
            for mode, th in [("nn", 0), ("mnn", 0), ("snn", 0.9), ("smnn", 0.95)]:
                idx, dist = ctx.nn_match(a, b, mode, th)
                ridx, rdist = o_nn.kornia_match({"descriptors": a}, {"descriptors": b}, mode, th)
                same = idx.shape == ridx.shape and np.array_equal(idx, ridx)
                dd = np.abs(dist - rdist).max() if same and len(dist) else -1
                print(f"  nn tc={tc} {mode} {n0}x{n1}: ours {len(idx)} ref {len(ridx)} identical={same} max|ddist|={dd:.2e}", flush=True)


def stage_time():
    import numpy as np
    import torch
    from dim_b200 import _native, synthetic, weights
    ctx = _native.Context(0)
    w = weights.superpoint_v1()
    conf = {"nms_radius": 3, "keypoint_threshold": 0.0005, "max_keypoints": 2048}
    B = 2
    sp = _native.SuperPointNet(ctx, w, max_batch=B, max_height=1024, max_width=1024, **conf)
    g = np.stack(synthetic.synthetic_pair(0, 1024))
    for prec in ("exact", "fast"):
        ctx.set_precision(prec)
        img = torch.from_numpy(g).cuda()
        kp = torch.zeros(B, 2048, 2, device="cuda"); sc = torch.zeros(B, 2048, device="cuda"); de = torch.zeros(B, 256, 2048, device="cuda")
        cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for it in range(3):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            sp.extract_dev(img.data_ptr(), B, 1024, 1024, kp.data_ptr(), sc.data_ptr(), de.data_ptr(), cnt.data_ptr(), 2048, st)
            e1.record(); torch.cuda.synchronize()
            print(f"  time sp {prec} B={B} 1024^2: {e0.elapsed_time(e1):.3f} ms  counts={cnt.tolist()}", flush=True)
        wl = weights.lightglue_seeded(seed=0)
        P = 2
        for name, dc, wc in [("fixed", -1, -1), ("adaptive", 0.95, 0.99)]:
            lg = _native.LightGlueNet(ctx, wl, depth_confidence=dc, width_confidence=wc, max_pairs=P, max_kpts=2048)
            m = torch.zeros(P, 2048, 2, dtype=torch.int64, device="cuda"); ms = torch.zeros(P, 2048, device="cuda")
            nm = torch.zeros(P, dtype=torch.int32, device="cuda"); sl = torch.zeros(P, dtype=torch.int32, device="cuda")
            fd = []
            for p in range(P):
                pair = []
                for s in range(2):
                    pair.append(_native.FeatsDev(kp[s].data_ptr(), de[s].data_ptr(), cnt[s:s + 1].data_ptr(), 2048, 0, 2048, 1024.0, 1024.0, 1))
                fd.append(pair)
            for it in range(3):
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record()
                lg.match_dev([f[0] for f in fd], [f[1] for f in fd], m.data_ptr(), ms.data_ptr(), nm.data_ptr(), sl.data_ptr(), 2048, st)
                e1.record(); torch.cuda.synchronize()
                print(f"  time lg {prec} {name} P={P} 2048x2048: {e0.elapsed_time(e1):.3f} ms  matches={nm.tolist()} stop={sl.tolist()}", flush=True)


def run_stage(name):
    if name.startswith("gemm"):
        stage_gemm(name.endswith("tc"))
    elif name == "sp_batch":
        stage_sp_batch()
    elif name.startswith("sp_"):
        stage_sp(name.endswith("tc"))
    elif name.startswith("lg_"):
        stage_lg(name.endswith("tc"))
    elif name == "nn":
        stage_nn()
    elif name == "time":
        stage_time()


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--stage":
        run_stage(sys.argv[2])
        sys.exit(0)
    stages = sys.argv[1:] or STAGES
    for s in stages:
        print(f"== stage {s}", flush=True)
        t = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--stage", s], timeout=300, capture_output=True, text=True)
            print(r.stdout[-6000:], flush=True)
            if r.returncode != 0:
                print(f"!! stage {s} exit {r.returncode}\n{r.stderr[-3000:]}", flush=True)
        except subprocess.TimeoutExpired as e:
            print(f"!! stage {s} TIMEOUT\n{(e.stdout or b'')[-3000:]}", flush=True)
        print(f"== stage {s} done in {time.time() - t:.1f}s", flush=True)
