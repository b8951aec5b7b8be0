"""Device feature store (the features.h5 boundary in HBM), the device-resident NN entry and the two-phase
image-set path - GPU parity against the host / oracle paths."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _feats(seed, n, d, hw):
    rng = np.random.default_rng(seed)
    desc = rng.standard_normal((d, n)).astype(np.float32)
    desc /= np.linalg.norm(desc, axis=0)
    return {"keypoints": (rng.uniform(0, 1, (n, 2)) * np.array([hw[1] - 1, hw[0] - 1])).astype(np.float32), "descriptors": desc,
            "scores": rng.uniform(0, 1, n).astype(np.float32), "tile_idx": rng.integers(0, 4, n).astype(np.float32),
            "image_size": np.array(hw)}


def test_feature_store_honours_the_h5_contract(ctx, tmp_path):
    """put -> get_features equals save_features_h5 -> get_features of the reference (value-level: float16 cast of every array,
    float32 / int32 on the way back, io/h5.py:45-89); unknown image -> ValueError; bulk write re-reads identically."""
    from dim_b200.io_h5 import FeatureStore, as_half_roundtrip, get_features
    store = FeatureStore(ctx, max_images=4, cap=3000, desc_dim=128)
    cases = {"a.jpg": _feats(1, 2500, 128, (1536, 2048)), "b.jpg": _feats(2, 0, 128, (480, 640)), "c.jpg": _feats(3, 17, 128, (101, 3001))}
    for name, f in cases.items():
        store.put(name, f)
    for name, f in cases.items():
        got, exp = store.get_features(name), as_half_roundtrip(f)
        for k in ("keypoints", "descriptors", "scores", "tile_idx", "image_size"):
            assert got[k].dtype == exp[k].dtype and got[k].shape == exp[k].shape and np.array_equal(got[k], exp[k]), (name, k)
    assert np.abs(store.get_features("a.jpg")["keypoints"] - cases["a.jpg"]["keypoints"]).max() <= 1.0  # fp16 quantisation above 1024 px
    with pytest.raises(ValueError, match="Cannot find image"):
        store.get_features("missing.jpg")
    with pytest.raises(TypeError):
        store.put("d.jpg", {"keypoints": [[0, 0]], "descriptors": np.zeros((128, 1), np.float32)})
    store.write_h5(tmp_path / "features.h5")
    for name, f in cases.items():
        back = get_features(tmp_path / "features.h5", name)
        for k, v in as_half_roundtrip(f).items():
            assert np.array_equal(back[k], v), (name, k)


def test_store_feeds_lightglue_and_nn_without_leaving_hbm(ctx, sp_weights):
    """extract_dev -> put_dev (float16 cast on device) -> match_dev straight from the store blocks == the host path extract ->
    as_half_roundtrip -> match; descriptors of the same blocks through dimb_nn_match_dev (fp16 input, single-MMA exact path) ==
    dimb_nn_match on the host copies."""
    import torch
    from dim_b200 import _native, synthetic, weights
    from dim_b200.io_h5 import FeatureStore, as_half_roundtrip
    size, K = 512, 1024
    conf = {"nms_radius": 3, "keypoint_threshold": 0.0005, "max_keypoints": K}
    g = np.stack(synthetic.synthetic_pair(19, size))
    sp = _native.SuperPointNet(ctx, sp_weights, max_batch=2, max_height=size, max_width=size, **conf)
    w = weights.lightglue_seeded(seed=0)
    lg = _native.LightGlueNet(ctx, w, max_pairs=1, max_kpts=K)
    store = FeatureStore(ctx, 2, K, 256)
    img = torch.from_numpy(g).cuda()
    kp = torch.zeros(2, K, 2, device="cuda"); sc = torch.zeros(2, K, device="cuda"); de = torch.zeros(2, 256, K, device="cuda")
    cnt = torch.zeros(2, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    sp.extract_dev(img.data_ptr(), 2, size, size, kp.data_ptr(), sc.data_ptr(), de.data_ptr(), cnt.data_ptr(), K, st)
    for s, name in enumerate(("l.jpg", "r.jpg")):
        store.put_dev(name, kp[s].data_ptr(), sc[s].data_ptr(), de[s].data_ptr(), K, cnt[s:s + 1].data_ptr(), size, size, st)
    m = torch.zeros(1, K, 2, dtype=torch.int64, device="cuda"); ms = torch.zeros(1, K, device="cuda")
    nm = torch.zeros(1, dtype=torch.int32, device="cuda"); sl = torch.zeros(1, dtype=torch.int32, device="cuda")
    lg.match_dev([store.feats_dev("l.jpg")], [store.feats_dev("r.jpg")], m.data_ptr(), ms.data_ptr(), nm.data_ptr(), sl.data_ptr(), K, st)
    torch.cuda.synchronize()
    host = [{"keypoints": kp[s, :int(cnt[s])].cpu().numpy(), "scores": sc[s, :int(cnt[s])].cpu().numpy(),
             "descriptors": de[s, :, :int(cnt[s])].cpu().numpy(), "image_size": np.array([size, size])} for s in range(2)]
    for s, name in enumerate(("l.jpg", "r.jpg")):  # the store holds exactly what the h5 round trip of the host copy gives
        got, exp = store.get_features(name), as_half_roundtrip(host[s])
        assert all(np.array_equal(got[k], exp[k]) for k in ("keypoints", "descriptors", "scores", "image_size"))
    f = [as_half_roundtrip(h) for h in host]
    ref = lg.match([({**f[0], "_layout": 0}, {**f[1], "_layout": 0})])[0]
    n = int(nm[0])
    assert n == len(ref["matches"]) > 100 and int(sl[0]) == ref["stop"]
    assert np.array_equal(m[0, :n].cpu().numpy(), ref["matches"]) and np.abs(ms[0, :n].cpu().numpy() - ref["scores"]).max() < 1e-6
    # ---- NN on the same blocks
    n0, n1 = int(cnt[0]), int(cnt[1])
    idx = torch.zeros(K, 2, dtype=torch.int64, device="cuda"); dist = torch.zeros(K, device="cuda"); dn = torch.zeros(1, dtype=torch.int32, device="cuda")
    for mode, th in (("mnn", 0.0), ("smnn", 0.9)):
        ctx.nn_match_dev(store.dev.desc_ptr(0), n0, store.dev.desc_ptr(1), n1, 256, mode, th, idx.data_ptr(), dist.data_ptr(), dn.data_ptr(),
                         K, f16=True, ld0=store.dev.cap, ld1=store.dev.cap, stream=st)
        torch.cuda.synchronize()
        ridx, rdist = ctx.nn_match(f[0]["descriptors"], f[1]["descriptors"], mode, th)
        k = int(dn[0])
        assert k == len(ridx) > 50 and np.array_equal(idx[:k].cpu().numpy(), ridx) and np.abs(dist[:k].cpu().numpy() - rdist).max() < 1e-6


@pytest.mark.parametrize("D", [32, 96, 100])
def test_nn_any_descriptor_size(ctx, D):
    """Descriptor sizes that are not a multiple of 64 (zero padded on device): kornia's DescriptorMatcher accepts any D."""
    from oracle import nn_match as o_nn
    rng = np.random.default_rng(D)
    a = rng.standard_normal((D, 300)).astype(np.float32)
    b = np.concatenate([a[:, :150] + 0.1 * rng.standard_normal((D, 150)).astype(np.float32), rng.standard_normal((D, 111)).astype(np.float32)], 1)
    for mode, th in (("mnn", 0.0), ("smnn", 0.9), ("snn", 0.8), ("nn", 0.0)):
        idx, dist = ctx.nn_match(a, b, mode, th)
        ridx, rdist = o_nn.kornia_match({"descriptors": a}, {"descriptors": b}, mode, th)
        assert np.array_equal(idx, ridx), (mode, len(idx), len(ridx))
        assert np.abs(dist - rdist).max() < 2e-5 * max(1.0, float(np.abs(rdist).max()))  # distances ~ sqrt(2 D): relative fp32 noise


def test_image_set_two_phase_equals_serial_plugins(ctx, sp_weights):
    """ImageSetMatcher (extract -> store -> exchange -> pair batches out of HBM; one process here, the NCCL exchange is covered by
    bench.py --mode exhaustive and tests/test_sharding_gloo.py) on 5 images / all 10 pairs == the reference-shaped serial loop
    through the plugins (SuperPointExtractor._extract, fp16 round trip, LightGlueMatcher._match_pairs per pair)."""
    import torch
    from dim_b200 import synthetic, weights
    from dim_b200.config import Config
    from dim_b200.extractors.superpoint import SuperPointExtractor
    from dim_b200.io_h5 import as_half_roundtrip
    from dim_b200.matchers.lightglue import LightGlueMatcher
    from dim_b200.pairs_generator import pairs_from_bruteforce
    from dim_b200.sharded import ImageSetMatcher
    size, K = 384, 512
    imgs = []
    for p in range(3):
        imgs += list(synthetic.synthetic_pair(70 + p, size))
    imgs = np.stack(imgs[:5]).astype(np.float32)
    sp_conf = {"nms_radius": 3, "keypoint_threshold": 0.0005, "max_keypoints": K}
    w = weights.lightglue_seeded(seed=0)
    pairs = pairs_from_bruteforce(list(range(5)))
    eng = ImageSetMatcher(ctx, sp_weights, w, 5, size, size, sp_conf, {}, batch_images=3, batch_pairs=4)
    tables = eng.run(torch.from_numpy(imgs).cuda(), list(range(5)), pairs)
    ext = SuperPointExtractor(Config(pipeline="superpoint+lightglue", extractor={"max_keypoints": K}))
    mat = LightGlueMatcher(Config(pipeline="superpoint+lightglue", matcher={"weights_dict": w}), local_features="superpoint")
    feats = []
    for g in imgs:
        f = ext._extract(g)
        f["image_size"] = np.array(g.shape[:2])
        feats.append(as_half_roundtrip(f))
    assert len(tables) == 10
    for (i, j), t in zip(pairs, tables):
        exp = mat._match_pairs(feats[i], feats[j])
        assert t.dtype == np.int64 and np.array_equal(t, exp), (i, j, len(t), len(exp))
    assert sum(len(t) for t in tables) > 300


def test_tile_preselection_and_match_by_tile(ctx, sp_weights):
    """tile_selection PRESELECTION (matcher_base.py:989-1148) - SuperPoint (nms 5, 4000 kpts, 0.005, hloc sampling)
    + LightGlue (0.9 / 0.95 / 0.3, no image_size) on the down-sampled images, matches scaled back, tile pairs with more than
    min_matches_per_tile common matches - then ExtractorBase._extract_by_tile + MatcherBase._match_by_tile over the selected
    pairs.  Against the oracle flows of oracle/tiling.py."""
    from dim_b200 import _native, synthetic, tiling, weights
    from dim_b200.config import Config
    from dim_b200.extractors.superpoint import SuperPointExtractor
    from dim_b200.io_h5 import as_half_roundtrip
    from dim_b200.matchers.lightglue import LightGlueMatcher
    from oracle import lightglue as o_lg
    from oracle import superpoint as o_sp
    from oracle import tiling as o_t
    a, b = synthetic.synthetic_pair(91, 1024)
    i0, i1 = a[:768].copy(), b[:768].copy()
    tile, ov, presel = (512, 512), 64, 512
    w_pre = weights.lightglue_seeded(seed=5)
    nets = {}

    def sp_factory(H, W):
        if (H, W) not in nets:
            nets[(H, W)] = _native.SuperPointNet(ctx, sp_weights, max_batch=1, max_height=H, max_width=W, **tiling.SP_PRESELECTION_CONF)
        return nets[(H, W)]

    lg_pre = _native.LightGlueNet(ctx, w_pre, max_pairs=1, max_kpts=4000, **tiling.LG_PRESELECTION_CONF)
    kp0, kp1 = tiling.preselection_matches(i0, i1, presel, sp_factory, lg_pre)
    ok0, ok1 = o_t.preselection_keypoints(i0, i1, presel, sp_weights, w_pre)
    assert len(kp0) == len(ok0) > 50 and np.array_equal(kp0, ok0) and np.array_equal(kp1, ok1)
    pairs = tiling.tile_selection(i0, i1, "preselection", tile, ov, kp0=kp0, kp1=kp1, min_matches_per_tile=5)
    assert pairs == o_t.select_tiles(i0, i1, "preselection", tile, ov, ok0, ok1, 5) and 1 <= len(pairs) <= 16
    assert tiling.tile_selection(i0, i1, "grid", tile, ov) == [(0, 0), (1, 1), (2, 2), (3, 3)]
    assert len(tiling.tile_selection(i0, i1, "exhaustive", tile, ov)) == 16
    # ---- extraction by tile (fix_sampling: with tiling enabled the reference's sampler is the patched one)
    ext_conf = {"max_keypoints": 512, "fix_sampling": True}
    cfg = Config(pipeline="superpoint+lightglue", extractor=ext_conf, general={"tile_size": tile, "tile_overlap": ov})
    ext = SuperPointExtractor(cfg)
    oconf = {**cfg.extractor}
    feats, ofeats = [], []
    for im in (i0, i1):
        f = ext._extract_by_tile(im)
        o = o_t.extract_by_tile(im, lambda t: o_sp.extract(np.ascontiguousarray(t[:, :, 0]), sp_weights, oconf), tile, ov, 256)
        assert np.array_equal(f["keypoints"], o["keypoints"]) and np.array_equal(f["tile_idx"], o["tile_idx"])
        assert np.abs(f["descriptors"] - o["descriptors"]).max() < TOL and np.abs(f["scores"] - o["scores"]).max() < TOL
        f["image_size"] = np.array(im.shape[:2])
        feats.append(as_half_roundtrip(f))
    # ---- matching by tile on exactly these features
    w = weights.lightglue_seeded(seed=0)
    m = LightGlueMatcher(Config(pipeline="superpoint+lightglue", matcher={"weights_dict": w}), local_features="superpoint")
    got = m._match_by_tile(feats[0], feats[1], pairs)
    exp = o_t.match_by_tile(feats[0], feats[1], pairs, lambda f0, f1: o_lg.match(f0, f1, w)["matches"])
    assert got.dtype == np.int64 and got.shape[1] == 2 and len(exp) > 30
    diff = {tuple(r) for r in got} ^ {tuple(r) for r in exp}
    assert len(diff) <= 2, diff  # a match at the 0.1 filter threshold of one tile pair may flip
