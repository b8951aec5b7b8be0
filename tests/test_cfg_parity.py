"""Parity of the CUDA path at the sizes BASELINE.json's configs state, on exactly what bench.py times (-m gpu).

cfg2: dimb_pipe_match_image_pairs{,_u8,_dev} - P pairs of 1024x1024 images, 2048 keypoints, two overlapped copy chunks, fp16
      rounding of the features on device - against the oracle chain SuperPoint -> features.h5 round trip -> LightGlue
      (reference image_matching.py:413-494, extractors/extractor_base.py:56-99), fixed-work and adaptive.
cfg3: ALIKED on a 2048x1536 image tiled 1024 / overlap 128 (4 tiles of 1024^2, 4096 keypoints each, nms 3) through
      ExtractorBase._extract_by_tile (extractor_base.py:279-390), and LightGlue at 4096 x 4096 keypoints, input_dim 128.
cfg5: brute-force NN at 8192 x 8192 x 256-d, mnn and smnn 0.99, index-exact against the oracle.
cfg1: the reference CPU plumbing case - OpenCV SIFT features of assets/example_sacre_coeur A / B (committed fixture) written to
      and re-read from the features store, matched with kornia_matcher smnn 0.85.
Tolerances: indices bit-exact, floats 1e-4 (north_star); where a test compares a CHAIN of two networks against the oracle
chain, fp16 roundings at the h5 boundary can flip on ~1e-6 descriptor differences and LightGlue amplifies them: those
comparisons are bounded by the oracle chain's own measured sensitivity to the same perturbation (and say so); the same call is
always also checked at 1e-4 on identical inputs.
"""
import os

import numpy as np
import pytest

from conftest import GOLD

pytestmark = pytest.mark.gpu
TOL = 1e-4
SIZE, KPTS, P = 1024, 2048, 4
SP_CONF = {"nms_radius": 3, "keypoint_threshold": 0.0005, "max_keypoints": KPTS}


@pytest.fixture(scope="module")
def cfg2(sp_weights):
    """P synthetic pairs of record + the oracle's features for every image (after the features.h5 round trip)."""
    from dim_b200 import synthetic
    from dim_b200.io_h5 import as_half_roundtrip
    from oracle import superpoint as o_sp
    # pairs of the generator of record whose top-2048 cut is decided by a clear margin in the oracle's own score map (gap between
    # the 2048th and 2049th candidate > 8e-5, several times the ~1e-5 score error of either implementation; about one pair in four qualifies): then every correct implementation selects the same keypoints.  (A tie at the cut
    # swaps one keypoint, which legitimately moves the LightGlue scores of every match competing with it - that case is covered by
    # tests/test_gpu_parity.py::test_superpoint_cfg2_full_size_batch, not by a chain comparison.)
    imgs, raw, seed = [], [], 40
    while len(imgs) < 2 * P and seed < 140:
        pair = synthetic.synthetic_pair(seed, SIZE)
        r = [o_sp.extract(g, sp_weights, SP_CONF, return_debug=True) for g in pair]
        ok = True
        for x in r:
            nms = x["_nms"][4:-4, 4:-4]
            cand = np.sort(nms[nms > SP_CONF["keypoint_threshold"]])[::-1]
            ok &= len(cand) > KPTS and float(cand[KPTS - 1] - cand[KPTS]) > 8e-5
        if ok:
            imgs += list(pair)
            raw += r
        seed += 1
    assert len(imgs) == 2 * P, "not enough synthetic pairs with a clear top-k margin"
    imgs = np.stack(imgs).astype(np.float32)
    feats = [as_half_roundtrip({"keypoints": r["keypoints"], "descriptors": r["descriptors"], "scores": r["scores"],
                                "image_size": np.array([SIZE, SIZE])}) for r in raw]
    return {"images": imgs, "raw": raw, "feats": feats}


def _pipe(ctx, sp_weights, fixed):
    from dim_b200 import _native, weights
    sp = _native.SuperPointNet(ctx, sp_weights, max_batch=2 * P, max_height=SIZE, max_width=SIZE, **SP_CONF)
    kw = {"depth_confidence": -1, "width_confidence": -1} if fixed else {}
    lg = _native.LightGlueNet(ctx, weights.lightglue_seeded(seed=0), max_pairs=P, max_kpts=KPTS, **kw)
    return _native.Pipe(sp, lg, P, SIZE, SIZE, KPTS), sp, lg


_LG_CACHE = {}


def _oracle_lg(f0, f1, w, conf):
    """oracle LightGlue, memoised on the exact input bytes (the three entry points produce bit-identical features)."""
    import hashlib
    from oracle import lightglue as o_lg
    h = hashlib.sha1()
    for f in (f0, f1):
        h.update(np.ascontiguousarray(f["keypoints"]).tobytes())
        h.update(np.ascontiguousarray(f["descriptors"]).tobytes())
    key = (h.hexdigest(), conf["depth_confidence"], conf["width_confidence"])
    if key not in _LG_CACHE:
        _LG_CACHE[key] = o_lg.match(f0, f1, w, conf)
    return _LG_CACHE[key]


_FLOOR_CACHE = {}


def _chain_noise_floor(cfg2, p, w_lg, conf, amp):
    """How far the ORACLE chain's match scores move when its own SuperPoint descriptors are perturbed by uniform noise of
    amplitude `amp` (the measured |GPU - oracle| descriptor error) before the float16 cast of features.h5: the resolution at which
    any two correct implementations of the chain can be expected to agree."""
    from dim_b200.io_h5 import as_half_roundtrip
    key = (p, conf["depth_confidence"], round(float(np.log10(amp)), 1))
    if key in _FLOOR_CACHE:
        return _FLOOR_CACHE[key]
    base = _oracle_lg(cfg2["feats"][2 * p], cfg2["feats"][2 * p + 1], w_lg, conf)
    mb = {tuple(m): s for m, s in zip(base["matches"], base["scores"])}
    rng = np.random.default_rng(123 + p)
    f = []
    for r in (cfg2["raw"][2 * p], cfg2["raw"][2 * p + 1]):
        d = r["descriptors"] + rng.uniform(-amp, amp, r["descriptors"].shape).astype(np.float32)
        f.append(as_half_roundtrip({"keypoints": r["keypoints"], "descriptors": d, "scores": r["scores"], "image_size": np.array([SIZE, SIZE])}))
    pert = _oracle_lg(f[0], f[1], w_lg, conf)
    mp = {tuple(m): s for m, s in zip(pert["matches"], pert["scores"])}
    _FLOOR_CACHE[key] = max([abs(mb[m] - mp[m]) for m in set(mb) & set(mp)] + [1e-4])
    return _FLOOR_CACHE[key]


def _as_coord_matches(m, k0, k1):
    return {(tuple(k0[i]), tuple(k1[j])) for i, j in m}


@pytest.mark.parametrize("entry", ["f32", "u8", "dev"])
@pytest.mark.parametrize("mode", ["fixed", "adaptive"])
def test_cfg2_pipe_is_what_the_oracle_chain_computes(ctx, sp_weights, cfg2, entry, mode):
    """The timed entry points of bench.py.  (1) every image's keypoints / scores / descriptors (read back from HBM) equal the
    oracle's SuperPoint at 1e-4; (2) LightGlue on EXACTLY the features the pipe produced (host fp16 round trip of the read-back
    arrays -> oracle LightGlue) gives identical matches / stop layer and scores within 1e-4; (3) against the pure oracle chain
    the match sets agree up to threshold-boundary cases and scores within 2e-3 (h5 rounding flips, see module docstring)."""
    import torch
    from dim_b200 import weights
    from dim_b200.io_h5 import as_half_roundtrip
    from oracle import lightglue as o_lg
    from oracle.compare import compare_matches, compare_superpoint
    fixed = mode == "fixed"
    pipe, sp, lg = _pipe(ctx, sp_weights, fixed)
    imgs = cfg2["images"]
    if entry == "f32":
        out = pipe.match_image_pairs(imgs, want_kpts=True)
    elif entry == "u8":
        assert np.array_equal(imgs, imgs.astype(np.uint8))  # synthetic gray images are integer valued
        out = pipe.match_image_pairs(imgs.astype(np.uint8), want_kpts=True)
    else:
        d_img = torch.from_numpy(imgs).cuda()
        pipe.match_image_pairs_dev(d_img.data_ptr(), P, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        o = pipe.outputs_dev()
        out = pipe.alloc_outputs(P, want_kpts=True)
        for key, name in (("matches", "matches"), ("mscores", "mscores"), ("n_matches", "n_matches"), ("stop", "stop"),
                          ("n_kpts", "n_kpts"), ("kpts", "kpts")):
            ctx.check(ctx.lib.dimb_read_dev(ctx.h, out[name].ctypes.data, o[key], out[name].nbytes), "dimb_read_dev")
    got_feats = pipe.read_features(P)
    w_lg = weights.lightglue_seeded(seed=0)
    conf = {**o_lg.DEFAULT_CONF, **({"depth_confidence": -1, "width_confidence": -1} if fixed else {})}
    worst, worst_desc = 0.0, 0.0
    cut = []  # per image: keypoints on which the two top-2048 selections differ (scores within 1e-4 of the cut, proven by compare_superpoint)
    for b in range(2 * P):  # (1) SuperPoint of the batched, chunked extraction
        assert out["n_kpts"][b] == len(got_feats[b]["keypoints"]) == KPTS
        rep = compare_superpoint(got_feats[b], cfg2["raw"][b], cfg2["raw"][b]["_nms"], TOL)
        cut.append({(float(x), float(y)) for x, y in rep["boundary_diffs"]})
        assert np.array_equal(out["kpts"][b, :KPTS], got_feats[b]["keypoints"])
        worst, worst_desc = max(worst, rep["max_dscore"], rep["max_ddesc"]), max(worst_desc, rep["max_ddesc"])
    for p in range(P):
        n = int(out["n_matches"][p])
        got = {"matches": out["matches"][p, :n], "scores": out["mscores"][p, :n], "stop": int(out["stop"][p])}
        f = [as_half_roundtrip({**got_feats[2 * p + s], "image_size": np.array([SIZE, SIZE])}) for s in (0, 1)]
        rep = compare_matches(got, _oracle_lg(f[0], f[1], w_lg, conf), 0.1, TOL)  # (2) same inputs: tight
        exp = _oracle_lg(cfg2["feats"][2 * p], cfg2["feats"][2 * p + 1], w_lg, conf)  # (3) the oracle chain
        assert got["stop"] == exp["stop"]
        a = _as_coord_matches(got["matches"], got_feats[2 * p]["keypoints"], got_feats[2 * p + 1]["keypoints"])
        e = _as_coord_matches(exp["matches"], cfg2["feats"][2 * p]["keypoints"], cfg2["feats"][2 * p + 1]["keypoints"])
        sc_e = {(tuple(cfg2["feats"][2 * p]["keypoints"][i]), tuple(cfg2["feats"][2 * p + 1]["keypoints"][j])): s
                for (i, j), s in zip(exp["matches"], exp["scores"])}
        sc_g = {(tuple(got_feats[2 * p]["keypoints"][i]), tuple(got_feats[2 * p + 1]["keypoints"][j])): s
                for (i, j), s in zip(got["matches"], got["scores"])}
        # (3) chain vs chain.  The reference's own chain is not reproducible below ~1e-3: descriptor differences of 1e-6 flip
        # ~7 % of the float16 roundings at the h5 boundary and LightGlue amplifies them (measured on the oracle alone, same
        # perturbation size as our descriptor error: score deltas 0.4e-3 .. 1.6e-3, see _chain_noise_floor below).  So: the
        # sets must agree except for low-confidence matches next to the 0.1 threshold, scores within 10x that floor.
        floor = _chain_noise_floor(cfg2, p, w_lg, conf, max(worst_desc, 1e-6))
        dchain = max((abs(sc_g[m] - sc_e[m]) for m in a & e), default=0.0)
        # a keypoint that only one of the two top-2048 selections holds (tie at the cut) takes its matches with it: not a
        # LightGlue difference.  Everything else may differ only next to the 0.1 filter threshold.
        on_cut = lambda m: (float(m[0][0]), float(m[0][1])) in cut[2 * p] or (float(m[1][0]), float(m[1][1])) in cut[2 * p + 1]
        rest = [m for m in a ^ e if not on_cut(m)]
        kg = ({tuple(k) for k in got_feats[2 * p]["keypoints"]}, {tuple(k) for k in got_feats[2 * p + 1]["keypoints"]})
        for m in rest:
            sc = sc_g.get(m, sc_e.get(m))
            assert sc < 0.1 + 20 * floor, (m, sc_g.get(m), sc_e.get(m), floor, "kp0 in GPU set", tuple(m[0]) in kg[0], "kp1 in GPU set",
                                           tuple(m[1]) in kg[1], "cut", sorted(cut[2 * p]), sorted(cut[2 * p + 1]))
        assert len(rest) <= max(4, len(e) // 100), (len(rest), len(e))
        assert dchain < max(10 * floor, 2e-3), (dchain, floor)
        print(f"{entry}/{mode} pair {p}: {n} matches, stop {got['stop']}, same-input dscore {rep['max_dscore']:.1e}, "
              f"chain dscore {dchain:.1e} (oracle-chain noise floor {floor:.1e}), set diff {len(rest)} (+{len(a ^ e) - len(rest)} on cut keypoints)")
    print(f"{entry}/{mode}: SuperPoint worst |delta| {worst:.1e} over {2 * P} images")


# ---------------------------------------------------------------------------------------------------- cfg3
@pytest.fixture(scope="module")
def cfg3_image():
    """(1536, 2048, 3) RGB: the block pattern of the generator of record at 4-px blocks (the n_limit = 4096 cut must fire)."""
    from dim_b200 import synthetic
    return synthetic.blocks_image(77, 2048, 4)[:1536].astype(np.float32)  # 4-px blocks: > 4096 ALIKED candidates per tile


def test_cfg3_aliked_tiled_2048x1536(ctx, al_weights, cfg3_image):
    """ExtractorBase._extract_by_tile with the AlikedExtractor plugin (tile 1024, overlap 128 -> 4 tiles incl. zero padding,
    max_num_keypoints 4096, nms 3) against the oracle flow: per-tile ALIKED at the full 1024^2 / 4096-keypoint size, keypoint
    shift, border mask, tile_idx, np.unique ordering."""
    from dim_b200.config import Config
    from dim_b200.extractors.aliked import AlikedExtractor
    from oracle import aliked as o_al
    from oracle import tiling as o_t
    from oracle.compare import compare_aliked
    ext_conf = {"max_num_keypoints": 4096}
    cfg = Config(pipeline="aliked+lightglue", extractor=ext_conf, general={"tile_size": (1024, 1024), "tile_overlap": 128})
    ext = AlikedExtractor(cfg)
    got = ext._extract_by_tile(cfg3_image)
    conf = {**cfg.extractor}
    per_tile = {}

    def oracle_extract(tile):
        f = o_al.extract(np.ascontiguousarray(tile), al_weights, conf, return_debug=True)
        per_tile[len(per_tile)] = f
        return f

    exp = o_t.extract_by_tile(cfg3_image, oracle_extract, (1024, 1024), 128, 128)
    assert len(per_tile) == 4 and max(len(f["keypoints"]) for f in per_tile.values()) == 4096  # the n_limit branch fires
    # per-tile parity at full size (the tile is what _extract sees: 1024 x 1024 x 3, zero padded top / bottom)
    tiles, _, _ = o_t.compute_tiles(cfg3_image, (1024, 1024), 128)
    worst = [0.0, 0.0, 0.0]
    for t in range(4):
        out = ext._extract(np.ascontiguousarray(tiles[t]))
        rep = compare_aliked(out, per_tile[t], per_tile[t]["_score_map"], conf["detection_threshold"], conf["nms_radius"], tol=TOL,
                             tol_kpt=1e-3, max_boundary=8)
        assert rep["n"] >= 4000
        worst = [max(a, b) for a, b in zip(worst, (rep["max_dkpt"], rep["max_dscore"], rep["max_ddesc"]))]
        print(f"tile {t}: {rep['n']} kpts, boundary {len(rep['boundary_diffs'])}, dkpt {rep['max_dkpt']:.1e} dscore {rep['max_dscore']:.1e} "
              f"ddesc {rep['max_ddesc']:.1e}")
    # the assembled FeaturesDict: every keypoint paired with the oracle's (nearest neighbour), same tile_idx / values, and the
    # np.unique(axis=0) order (lexicographic by x, then y)
    from scipy.spatial import cKDTree
    assert got["keypoints"].shape[0] == got["tile_idx"].shape[0] == got["descriptors"].shape[1] == got["scores"].shape[0]
    assert abs(len(got["keypoints"]) - len(exp["keypoints"])) <= 16
    dist, j = cKDTree(exp["keypoints"].astype(np.float64)).query(got["keypoints"].astype(np.float64))
    ok = dist < 1e-3
    assert ok.mean() > 0.995, ok.mean()
    assert np.array_equal(got["tile_idx"][ok], exp["tile_idx"][j[ok]])
    assert np.abs(got["descriptors"][:, ok] - exp["descriptors"][:, j[ok]]).max() < TOL
    assert np.abs(got["scores"][ok] - exp["scores"][j[ok]]).max() < TOL
    assert np.all(np.diff(got["keypoints"][:, 0]) >= 0)
    print("assembled:", len(got["keypoints"]), "keypoints (oracle", len(exp["keypoints"]), "), worst per tile", worst)


def test_cfg3_lightglue_4096x4096_input_dim_128(ctx):
    """LightGlue at the cfg3 tile-pair size: 4096 x 4096 keypoints, 128-d descriptors (input projection), 9 layers."""
    from dim_b200 import _native
    from oracle import lightglue as o_lg
    from oracle.compare import compare_matches
    from oracle.gen_golden import lg_pair
    f0, f1 = lg_pair(61, 4096, 4096, 128, (1536, 2048))
    for name, over in (("fixed", {"depth_confidence": -1, "width_confidence": -1}), ("adaptive", {})):
        conf = {**o_lg.DEFAULT_CONF, "input_dim": 128, **over}
        w = o_lg.seeded_weights(conf, seed=61)
        lg = _native.LightGlueNet(ctx, w, input_dim=128, depth_confidence=conf["depth_confidence"],
                                  width_confidence=conf["width_confidence"], max_pairs=1, max_kpts=4096)
        out = lg.match([({**f0, "_layout": 0}, {**f1, "_layout": 0})])[0]
        exp = o_lg.match(f0, f1, w, conf)
        rep = compare_matches(out, exp, 0.1, TOL)
        print(name, rep["n"], "matches, stop", out["stop"], "max score delta", rep["max_dscore"], rep["boundary_diffs"])
        assert rep["n"] > 1500
        del lg


# ---------------------------------------------------------------------------------------------------- cfg5
def _cfg5_descriptors(seed, n=8192, d=256):
    """unit-norm gaussian descriptors rounded to fp16 (as read from features.h5)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((d, n)).astype(np.float32)
    x /= np.linalg.norm(x, axis=0)
    return x.astype(np.float16).astype(np.float32)


@pytest.mark.parametrize("mode,th", [("mnn", 0.0), ("smnn", 0.99)])
def test_cfg5_nn_8192x256_index_exact(ctx, mode, th):
    """Unrelated random descriptor sets (the hard case: nearest / second-nearest distances ~1 % apart, ratios around the 0.99
    threshold) plus a block of true correspondences.  Index pairs must equal the oracle's; a row may differ only where the
    oracle's own decision margin (top-2 gap, or ratio - th) is below the fp32 noise of a 256-term dot product (1e-6), and
    every such row is printed."""
    import torch
    from oracle import nn_match as o_nn
    d0, d1 = _cfg5_descriptors(5), _cfg5_descriptors(6)
    rng = np.random.default_rng(7)
    src = rng.permutation(8192)[:2048]
    mix = d0[:, src] + 0.05 * rng.standard_normal((256, 2048)).astype(np.float32)
    d1[:, rng.permutation(8192)[:2048]] = (mix / np.linalg.norm(mix, axis=0)).astype(np.float16).astype(np.float32)
    idx, dist = ctx.nn_match(d0, d1, mode, th)
    ridx, rdist = o_nn.kornia_match({"descriptors": d0}, {"descriptors": d1}, mode, th)
    got, exp = {tuple(r) for r in idx}, {tuple(r) for r in ridx}
    if got != exp:
        dm = torch.cdist(torch.from_numpy(d0.T.copy()), torch.from_numpy(d1.T.copy()))
        v01, _ = torch.topk(dm, 2, dim=1, largest=False)
        v10, _ = torch.topk(dm.t(), 2, dim=1, largest=False)
        for i, j in sorted(got ^ exp):
            gap = min(float(v01[i, 1] - v01[i, 0]), float(v10[j, 1] - v10[j, 0]))
            near_th = min(abs(float(v01[i, 0] / v01[i, 1]) - th), abs(float(v10[j, 0] / v10[j, 1]) - th)) if mode == "smnn" else 1.0
            print("differs:", (i, j), "top-2 gap", gap, "ratio margin", near_th)
            assert gap < 1e-6 or near_th < 1e-6
    assert len(got ^ exp) <= 2 and len(idx) > 1500
    if len(got ^ exp) == 0:
        assert np.array_equal(idx, ridx) and np.abs(dist - rdist).max() < TOL
    print(mode, len(idx), "matches of 8192 x 8192")


# ---------------------------------------------------------------------------------------------------- cfg1
def test_cfg1_sift_kornia_matcher_plumbing(ctx, tmp_path):
    """The reference's CPU pipeline sift + kornia_matcher (config.py:234-244) on assets/example_sacre_coeur A / B: the SIFT
    features (fixture generated with OpenCV exactly as SIFTExtractor._extract does, oracle/gen_golden.py:gen_cfg1) are written
    through save_features_h5, re-read through get_features (fp16 round trip, io/h5.py:45-89) and matched by the KorniaMatcher
    plugin (smnn 0.85) -> the oracle's match table; mnn as a second mode."""
    from dim_b200.config import Config
    from dim_b200.io_h5 import get_features, save_features_h5
    from dim_b200.matchers.kornia_matcher import KorniaMatcher
    g = np.load(os.path.join(GOLD, "cfg1_sift_golden.npz"))
    path = tmp_path / "features.h5"
    for tag, name in (("0", "sacre_coeur_A.jpg"), ("1", "sacre_coeur_B.jpg")):
        feats = {"keypoints": g["kpts" + tag], "descriptors": g["desc" + tag].astype(np.float64),  # SIFTExtractor: des.astype(float).T
                 "tile_idx": np.zeros(len(g["kpts" + tag]), np.float32), "image_size": g["size" + tag]}
        save_features_h5(path, feats, name)
    f0, f1 = get_features(path, "sacre_coeur_A.jpg"), get_features(path, "sacre_coeur_B.jpg")
    assert f0["descriptors"].dtype == np.float32 and f0["descriptors"].shape == (128, 2048) and f1["descriptors"].shape == (128, 2049)
    assert np.array_equal(f0["descriptors"], g["desc0"].astype(np.float32)) and np.array_equal(f0["image_size"], g["size0"])
    assert np.array_equal(f0["keypoints"], g["kpts0"].astype(np.float16).astype(np.float32))
    m = KorniaMatcher(Config(pipeline="sift+kornia_matcher"))
    got = m._match_pairs(f0, f1)
    assert got.dtype == np.int64 and np.array_equal(got, g["smnn085.matches"].astype(np.int64)) and len(got) == 92
    m2 = KorniaMatcher(Config(matcher={"name": "kornia_matcher", "match_mode": "mnn"}))
    assert np.array_equal(m2._match_pairs(f0, f1), g["mnn.matches"].astype(np.int64))
