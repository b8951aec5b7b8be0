"""Measurements of the widened rows of SURVEY 8(d) (not the headline - that is bench.py):

  cfg3a  ALIKED-n16rot extraction of one 1024x1024 RGB tile, max 4096 keypoints, nms 3, threshold 0.2   (HBM roofline)
  cfg3b  LightGlue on ALIKED-shaped features: input_dim 128, 4096 x 4096 keypoints, fixed work         (tensor roofline)
  cfg5   KorniaMatcher: 8192 x 8192 x 256 descriptors, smnn th 0.99 and mnn                             (tensor roofline)

Prints one JSON line per workload and writes them to gpurun_out/widened.json.  Timing: CUDA events around the call of
the C-ABI entry (host buffers for cfg3a / cfg5 as the plugins call them, device-resident for cfg3b), >= 3 warm-ups.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    return d


def rgb_tile(seed, size=1024):
    """blocks-16 generator of record in RGB."""
    import cv2
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, (size // 16, size // 16, 3), np.uint8)
    img = cv2.resize(small, (size, size), interpolation=cv2.INTER_NEAREST)
    return cv2.GaussianBlur(img, (0, 0), 0.8).astype(np.float32)


def timed(fn, steps, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch
    from dim_b200 import _native, weights
    ctx = _native.Context(0)
    pk = peaks()
    hbm_peak = pk.get("hbm_gbs") or 6485.0
    tf_peak = pk.get("bf16_tflops_sustained") or 1422.7
    out = []

    if not args.only or "plugin" in args.only:
        # what a drop-in user of the reference sees: the batch-1 plugin calls of image_matching.py:413-494
        from dim_b200 import synthetic
        from dim_b200.config import Config
        from dim_b200.extractors.superpoint import SuperPointExtractor
        from dim_b200.io_h5 import as_half_roundtrip
        from dim_b200.matchers.lightglue import LightGlueMatcher
        cfg = Config(pipeline="superpoint+lightglue")
        ext = SuperPointExtractor(cfg)
        mat = LightGlueMatcher(Config(pipeline="superpoint+lightglue", matcher={"weights_dict": weights.lightglue_seeded(seed=0)}))
        g0, g1 = synthetic.synthetic_pair(0, 1024)
        f = [as_half_roundtrip({**ext._extract(g), "image_size": np.array([1024, 1024])}) for g in (g0, g1)]
        nm = len(mat._match_pairs(f[0], f[1]))
        ms_ext = timed(lambda: ext._extract(g0), args.steps)
        ms_mat = timed(lambda: mat._match_pairs(f[0], f[1]), args.steps)
        out.append({"workload": "cfg2 through the batch-1 plugin calls (SuperPointExtractor._extract on one 1024x1024 image, "
                                "LightGlueMatcher._match_pairs on one 2048x2048 pair, adaptive defaults, host numpy in/out)",
                    "ms_extract_per_image": ms_ext, "ms_match_per_pair": ms_mat, "pairs_per_s_independent": 1e3 / (2 * ms_ext + ms_mat),
                    "n_matches": nm})
        print(json.dumps(out[-1]), flush=True)
        del ext, mat

    if not args.only or "aliked" in args.only:
        S = 1024
        net = _native.AlikedNet(ctx, weights.aliked_n16rot(), 4096, 0.2, 3, S, S)
        imgs = [rgb_tile(s, S) for s in range(3)]
        n = [len(net.extract(i)["keypoints"]) for i in imgs]
        # device-resident timing through dimb_aliked_extract_dev
        d_imgs = [torch.from_numpy(i).cuda() for i in imgs]
        cap = 4096
        kp = torch.zeros(cap, 2, device="cuda"); sc = torch.zeros(cap, device="cuda"); de = torch.zeros(128, cap, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        k = [0]

        def dev():
            net.extract_dev(d_imgs[k[0] % 3].data_ptr(), S, S, 3, kp.data_ptr(), sc.data_ptr(), de.data_ptr(), cnt.data_ptr(), cap, st)
            k[0] += 1

        def host():
            net.extract(imgs[k[0] % 3]); k[0] += 1
        ms_dev = timed(dev, args.steps)
        ctx.profile(True)
        for _ in range(args.steps):
            dev()
        torch.cuda.synchronize()
        prof = ctx.profile_read()
        ctx.profile(False)
        ms_host = timed(host, args.steps)
        alg_bytes = 12.6e6 + 70e6            # input + fp16 pyramid (algorithmic minimum per tile)
        # bytes this fp32 implementation must move by design: input, fp32 pyramid (x1 16@P w+r twice, x2.., ) and the
        # fp32 128@P feature map written once
        design_bytes = 4.0 * S * S * (3 + 3 + 16 * 2 + 16 * 2 + 16 + 8 * 2 + 4 * 4 + 2 + 128) + 4.0 * S * S / 4 * (16 + 32 * 6)
        out.append({"workload": "cfg3a: ALIKED-n16rot, one 1024x1024 RGB tile, max 4096 kpts, nms 3", "metric": "tiles/s",
                    "value": 1e3 / ms_dev, "ms_per_tile": ms_dev, "e2e_host_buffers": {"value": 1e3 / ms_host, "ms": ms_host},
                    "n_kpts": n, "dtype": "f32", "kernel_groups_ms_per_tile": {g: v[0] / args.steps for g, v in prof.items()},
                    "roofline": {"bound": "hbm", "achieved": alg_bytes / ms_dev / 1e6, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / ms_dev / 1e6 / hbm_peak,
                                 "design_bytes_gbps": design_bytes / ms_dev / 1e6,
                                 "note": "achieved = SURVEY 8(d) algorithmic bytes (82.6 MB / tile) / time; design_bytes = what "
                                         "the fp32 planar implementation moves by construction (feature map 537 MB written once)"}})
        print(json.dumps(out[-1]), flush=True)
        del net

    if not args.only or "lg" in args.only:
        def lg_pair(seed, m, n, dim, size_hw, overlap=0.6, noise=0.05):
            """two seeded feature sets sharing ~overlap*min(m,n) noisy, permuted correspondences (fp16-representable)"""
            rng = np.random.default_rng(seed)

            def feats(k):
                d = rng.standard_normal((dim, k)).astype(np.float32)
                return {"keypoints": (rng.random((k, 2)) * (min(size_hw) - 1)).astype(np.float32),
                        "descriptors": d / np.linalg.norm(d, axis=0, keepdims=True), "image_size": np.array(size_hw, np.int32)}
            f0, f1 = feats(m), feats(n)
            k = int(overlap * min(m, n))
            src, dst = rng.permutation(m)[:k], rng.permutation(n)[:k]
            f1["keypoints"][dst] = np.clip(f0["keypoints"][src] + rng.normal(0, 2.0, (k, 2)), 0, min(size_hw) - 1).astype(np.float32)
            d = f0["descriptors"][:, src] + noise * rng.standard_normal((dim, k)).astype(np.float32)
            f1["descriptors"][:, dst] = d / np.linalg.norm(d, axis=0, keepdims=True)
            for f in (f0, f1):
                f["keypoints"] = f["keypoints"].astype(np.float16).astype(np.float32)
                f["descriptors"] = f["descriptors"].astype(np.float16).astype(np.float32)
                f["_layout"] = 0  # (D,N) FeaturesDict layout
            return f0, f1
        N = 4096
        w = weights.lightglue_seeded(input_dim=128, seed=1)
        P = 4
        lg = _native.LightGlueNet(ctx, w, input_dim=128, depth_confidence=-1, width_confidence=-1, max_pairs=P, max_kpts=N)
        pairs = [lg_pair(10 + p, N, N, 128, (1024, 1024)) for p in range(P)]
        res = lg.match(pairs)
        ms = timed(lambda: lg.match(pairs), max(args.steps // 2, 3))
        gflop = 812.3 * P
        out.append({"workload": f"cfg3b: LightGlue input_dim 128, {N}x{N} kpts, fixed work, {P} tile pairs per call (host feature buffers)",
                    "metric": "tile-pairs/s", "value": P * 1e3 / ms, "ms_per_call": ms, "dtype": "f16 hi/lo split x3 MMA, f32 accumulate",
                    "n_matches": [int(len(r["matches"])) for r in res],
                    "roofline": {"bound": "tensor", "achieved": gflop / ms, "peak": tf_peak, "unit": "TFLOP/s", "frac": gflop / ms / tf_peak,
                                 "note": "algorithmic 812.3 GFLOP per tile pair (SURVEY 8d); includes H2D of descriptors"}})
        print(json.dumps(out[-1]), flush=True)
        del lg

    if not args.only or "lighterglue" in args.only:
        g = np.load(os.path.join(ROOT, "tests", "golden", "lighterglue_golden.npz"))
        w = weights.load_npz(os.path.join(weights.DATA, "lighterglue_weights.npz"))
        f = [{"keypoints": g[f"kpts{i}"].astype(np.float32), "descriptors": g[f"desc{i}"].astype(np.float32), "image_size": g[f"size{i}"],
              "_layout": 0} for i in (0, 1)]
        net = _native.LightGlueNet(ctx, w, input_dim=64, descriptor_dim=96, n_layers=6, num_heads=1, depth_confidence=-1,
                                   width_confidence=0.95, max_pairs=1, max_kpts=2048)
        r = net.match([(f[0], f[1])])[0]
        ms = timed(lambda: net.match([(f[0], f[1])]), args.steps)
        tc = os.environ.get("DIMB_TC", "1") != "0"
        out.append({"workload": "LighterGlue (trained weights, d 96 / 1 head / 6 layers), 2048 x 2048 XFeat keypoints, shape-generic path: "
                                + ("attention on the tensor cores (attn_hd128.cuh), fp32 linears, two side streams" if tc else "all fp32 (DIMB_TC=0)"),
                    "metric": "pairs/s", "value": 1e3 / ms, "ms_per_pair": ms, "n_matches": int(len(r["matches"])), "stop": r["stop"],
                    "dtype": "see workload"})
        print(json.dumps(out[-1]), flush=True)
        del net

    if not args.only or "superglue" in args.only:
        rng = np.random.default_rng(5)
        n = 2048

        def sg_feats():
            d = rng.standard_normal((256, n)).astype(np.float32)
            return {"keypoints": (rng.random((n, 2)) * 1023).astype(np.float32), "descriptors": d / np.linalg.norm(d, axis=0, keepdims=True),
                    "scores": rng.random(n).astype(np.float32), "image_size": np.array([1024, 1024])}
        names = _native.superglue_weight_names()
        shapes = {}
        ch = [3, 32, 64, 128, 256, 256]
        wsg = {}
        for nm in names:  # timing only: random weights of the right shapes
            if nm == "bin_score":
                wsg[nm] = np.array(1.0, np.float32)
        sgn = None
        try:
            for i in range(5):
                wsg[f"kenc.encoder.{3 * i}.weight"] = (rng.standard_normal((ch[i + 1], ch[i], 1)) / np.sqrt(ch[i])).astype(np.float32)
                wsg[f"kenc.encoder.{3 * i}.bias"] = np.zeros(ch[i + 1], np.float32)
                if i < 4:
                    for sfx, val in ((".weight", 1.0), (".bias", 0.0), (".running_mean", 0.0), (".running_var", 1.0)):
                        wsg[f"kenc.encoder.{3 * i + 1}{sfx}"] = np.full(ch[i + 1], val, np.float32)
            for i in range(18):
                p_ = f"gnn.layers.{i}."
                for nm, (co, ci) in (("attn.merge", (256, 256)), ("attn.proj.0", (256, 256)), ("attn.proj.1", (256, 256)), ("attn.proj.2", (256, 256)),
                                     ("mlp.0", (512, 512)), ("mlp.3", (256, 512))):
                    wsg[p_ + nm + ".weight"] = (0.5 * rng.standard_normal((co, ci, 1)) / np.sqrt(ci)).astype(np.float32)
                    wsg[p_ + nm + ".bias"] = np.zeros(co, np.float32)
                for sfx, val in ((".weight", 1.0), (".bias", 0.0), (".running_mean", 0.0), (".running_var", 1.0)):
                    wsg[p_ + "mlp.1" + sfx] = np.full(512, val, np.float32)
            wsg["final_proj.weight"] = (rng.standard_normal((256, 256, 1)) / 16).astype(np.float32)
            wsg["final_proj.bias"] = np.zeros(256, np.float32)
            sgn = _native.SuperGlueNet(ctx, wsg, max_kpts=n)
            fa, fb = sg_feats(), sg_feats()
            r = sgn.match(fa, fb)
            ms = timed(lambda: sgn.match(fa, fb), max(args.steps // 2, 3))
            out.append({"workload": "SuperGlue 2048 x 2048 keypoints, 18 layers, 100 Sinkhorn iterations, tensor-core path (random weights: timing only)",
                        "metric": "pairs/s", "value": 1e3 / ms, "ms_per_pair": ms, "n_matches": int(len(r["matches"])), "dtype": "see workload"})
            print(json.dumps(out[-1]), flush=True)
        finally:
            del sgn

    if not args.only or "nn" in args.only:
        rng = np.random.default_rng(0)
        n = 8192

        def desc():
            d = rng.standard_normal((256, n)).astype(np.float32)
            d /= np.linalg.norm(d, axis=0, keepdims=True)
            return d.astype(np.float16).astype(np.float32)
        a = desc()
        b = np.concatenate([a[:, : n // 2] + 0.02 * rng.standard_normal((256, n // 2)).astype(np.float32), desc()[:, n // 2:]], 1)
        b = (b / np.linalg.norm(b, axis=0, keepdims=True)).astype(np.float16).astype(np.float32)
        for mode, th in (("smnn", 0.99), ("mnn", 0.0)):
            m, _ = ctx.nn_match(a, b, mode, th)
            ms = timed(lambda: ctx.nn_match(a, b, mode, th), args.steps)
            ctx.profile(True)
            for _ in range(args.steps):
                ctx.nn_match(a, b, mode, th)
            prof = ctx.profile_read()
            ctx.profile(False)
            out.append({"workload": f"cfg5: kornia_matcher {mode} th {th}, 8192x8192x256 (host descriptor buffers)", "metric": "pairs/s",
                        "value": 1e3 / ms, "ms_per_pair": ms, "n_matches": int(len(m)), "dtype": "f16 hi/lo split x3 MMA, f32 accumulate",
                        "kernel_groups_ms": {g: v[0] / args.steps for g, v in prof.items()},
                        "roofline": {"bound": "tensor", "achieved": 34.4 / ms, "peak": tf_peak, "unit": "TFLOP/s", "frac": 34.4 / ms / tf_peak,
                                     "note": "algorithmic 34.4 GFLOP per pair; the timed call includes H2D of 16.8 MB descriptors and D2H of matches"}})
            print(json.dumps(out[-1]), flush=True)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "widened.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
