"""Generate tests/golden/*.npz by executing the REFERENCE's own model code.

Runs only in the authoring container (needs /root/reference).  It imports the
vendored model files *by path* (bypassing package __init__s that need
kornia/h5py), runs them on seeded inputs, checks that this repo's oracle agrees
and stores small fixtures so the agreement is re-checked wherever the tests
run (the GPU box has no /root/reference).

    python oracle/gen_golden.py            # writes tests/golden/
"""
from __future__ import annotations

import importlib.util
import os
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/src/deep_image_matching/"
T = REF + "thirdparty/"
GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "deep-image-matching_b200", "data")  # converted checkpoints shipped with the product package

from oracle import lightglue as o_lg  # noqa: E402
from oracle import nn_match as o_nn  # noqa: E402
from oracle import superpoint as o_sp  # noqa: E402
from dim_b200 import synthetic  # noqa: E402


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def sp_weights():
    sd = torch.load(T + "SuperGluePretrainedNetwork/models/weights/superpoint_v1.pth", map_location="cpu")
    return sd, {k: v.numpy().astype(np.float32) for k, v in sd.items()}


def fix_sampling_variant():
    """The patched sampler of extractors/superpoint.py:16-27, taken from that file's source."""
    src = open(REF + "extractors/superpoint.py").read()
    start = src.index("def sample_descriptors_fix_sampling")
    end = src.index("class SuperPoint(nn.Module)")
    ns = {"torch": torch}
    exec(src[start:end], ns)
    return ns["sample_descriptors_fix_sampling"]


def run_ref_sp(spmod, image, conf, fix):
    orig = spmod.sample_descriptors
    if fix:
        spmod.sample_descriptors = fix_sampling_variant()
    try:
        net = spmod.SuperPoint(conf).eval()
        with torch.no_grad():
            x = torch.tensor(image / 255.0, dtype=torch.float)[None, None]
            out = net({"image": x})
    finally:
        spmod.sample_descriptors = orig
    return {k: v[0].numpy() for k, v in out.items()}


def check_sp(name, ref, ora):
    po, pr = o_sp.canonical_order(ora), o_sp.canonical_order(ref)
    assert ref["keypoints"].shape == ora["keypoints"].shape, (name, ref["keypoints"].shape, ora["keypoints"].shape)
    assert np.array_equal(ref["keypoints"][pr], ora["keypoints"][po]), name
    ds = np.abs(ref["scores"][pr] - ora["scores"][po]).max()
    dd = np.abs(ref["descriptors"][:, pr] - ora["descriptors"][:, po]).max()
    print(f"  [{name}] N={len(po)} max|dscore|={ds:.2e} max|ddesc|={dd:.2e}")
    assert ds < 1e-5 and dd < 1e-5


def gen_superpoint():
    sd, w = sp_weights()
    np.savez(os.path.join(DATA, "superpoint_v1_weights.npz"), **w)
    torch.hub.load_state_dict_from_url = lambda *a, **k: sd
    spmod = load_by_path("ref_superpoint", T + "SuperGluePretrainedNetwork/models/superpoint.py")

    cases = {}
    # (a) real photograph from the reference's own test assets, cropped (keeps the fixture small)
    rgb = cv2.cvtColor(cv2.imread("/root/reference/assets/pytest/images/DSC_6466.jpg"), cv2.COLOR_BGR2RGB)
    real = synthetic.to_gray_like_reference(rgb)[100:340, 200:520].copy()  # 240x320
    cases["real240x320"] = (real, {"nms_radius": 3, "keypoint_threshold": 0.0005, "max_keypoints": 2048}, False)
    cases["real240x320_fix_top256"] = (real, {"nms_radius": 4, "keypoint_threshold": 0.005, "max_keypoints": 256}, True)
    # (b) synthetic generator of record, small and non-square
    g = synthetic.to_gray_like_reference(synthetic.blocks_image(3, 512))[:384, :]
    cases["blocks384x512_top512"] = (g, {"nms_radius": 3, "keypoint_threshold": 0.0005, "max_keypoints": 512}, False)
    # (c) H, W not multiples of 8 (pooling floors)
    cases["real_odd237x315"] = (real[:237, :315].copy(), {"nms_radius": 2, "keypoint_threshold": 0.001, "max_keypoints": -1}, False)
    out = {}
    for name, (img, conf, fix) in cases.items():
        ref = run_ref_sp(spmod, img, conf, fix)
        ora = o_sp.extract(img, w, {**conf, "fix_sampling": fix})
        check_sp(name, ref, ora)
        out[name + ".image"] = img.astype(np.uint8)
        out[name + ".conf"] = np.array([conf["nms_radius"], conf["keypoint_threshold"], conf["max_keypoints"], int(fix)], np.float64)
        out[name + ".keypoints"] = ref["keypoints"].astype(np.int16)
        out[name + ".scores"] = ref["scores"]
        out[name + ".descriptors"] = ref["descriptors"].astype(np.float32)
    # (d) the full-size BASELINE config-2 image: keep keypoints+scores and a descriptor subsample
    g0, _ = synthetic.synthetic_pair(0, 1024)
    conf = {"nms_radius": 3, "keypoint_threshold": 0.0005, "max_keypoints": 2048}
    ref = run_ref_sp(spmod, g0, conf, False)
    ora = o_sp.extract(g0, w, conf)
    check_sp("cfg2_pair0_img0", ref, ora)
    pr = o_sp.canonical_order(ref)
    out["cfg2.keypoints"] = ref["keypoints"][pr].astype(np.int16)
    out["cfg2.scores"] = ref["scores"][pr]
    out["cfg2.descriptors_first64"] = ref["descriptors"][:, pr[:64]].astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "superpoint_golden.npz"), **out)


def make_lg_features(rng, n, dim, size_hw, fp16=True):
    """Features as they arrive from features.h5: unit descriptors rounded to fp16, (D,N) layout."""
    h, w = size_hw
    kp = np.stack([rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n)], 1).astype(np.float32)
    d = rng.standard_normal((dim, n)).astype(np.float32)
    d /= np.linalg.norm(d, axis=0, keepdims=True)
    if fp16:
        kp = kp.astype(np.float16).astype(np.float32)
        d = d.astype(np.float16).astype(np.float32)
    return {"keypoints": kp, "descriptors": d, "scores": rng.uniform(0, 1, n).astype(np.float32),
            "tile_idx": np.zeros(n, np.float32), "image_size": np.array(size_hw, np.int32)}


def lg_pair(seed, m, n, dim, size_hw, overlap=0.6, noise=0.05):
    """Two feature sets sharing ~overlap*min(m,n) true correspondences (permuted, noisy)."""
    rng = np.random.default_rng(seed)
    f0 = make_lg_features(rng, m, dim, size_hw, fp16=False)
    f1 = make_lg_features(rng, n, dim, size_hw, fp16=False)
    k = int(overlap * min(m, n))
    src = rng.permutation(m)[:k]
    dst = rng.permutation(n)[:k]
    f1["keypoints"][dst] = np.clip(f0["keypoints"][src] + rng.normal(0, 2.0, (k, 2)), 0, min(size_hw) - 1).astype(np.float32)
    d = f0["descriptors"][:, src] + noise * rng.standard_normal((dim, k)).astype(np.float32)
    f1["descriptors"][:, dst] = d / np.linalg.norm(d, axis=0, keepdims=True)
    for f in (f0, f1):
        f["keypoints"] = f["keypoints"].astype(np.float16).astype(np.float32)
        f["descriptors"] = f["descriptors"].astype(np.float16).astype(np.float32)
    return f0, f1


def run_ref_lg(lgmod, w, conf, f0, f1):
    """Drive the reference LightGlue class exactly as LightGlueMatcher._match_pairs does (CPU)."""
    lgmod.LightGlue.pruning_keypoint_thresholds["cpu"] = conf.get("prune_min_kpts", 1536)
    net = lgmod.LightGlue(features=None, input_dim=conf["input_dim"], descriptor_dim=conf["descriptor_dim"],
                          n_layers=conf["n_layers"], num_heads=conf["num_heads"], flash=False,
                          depth_confidence=conf["depth_confidence"], width_confidence=conf["width_confidence"],
                          filter_threshold=conf["filter_threshold"]).eval()
    missing = net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    assert missing.missing_keys == ["confidence_thresholds"] or not missing.missing_keys, missing
    assert not missing.unexpected_keys, missing

    def conv(feats):  # featuresDict2Lightglue semantics (matchers/lightglue.py:8-66)
        k, d, s = o_lg.features_to_lg(feats)
        out = {"keypoints": k[None], "descriptors": d[None]}
        for key in ("scores", "tile_idx", "image_size"):
            if key in feats:
                out[key] = torch.as_tensor(np.asarray(feats[key]), dtype=torch.float32)[None]
        return out

    with torch.no_grad():
        res = net({"image0": conv(f0), "image1": conv(f1)})
    return {"matches": res["matches"][0].numpy(), "scores": res["scores"][0].numpy(), "stop": int(res["stop"]),
            "prune0": res["prune0"][0].numpy(), "prune1": res["prune1"][0].numpy()}


def gen_lightglue():
    lgmod = load_by_path("ref_lightglue", T + "LightGlue/lightglue/lightglue.py")
    out = {}
    cases = [
        # name, seed, m, n, conf overrides, size_hw
        ("sp_small_fixed", 1, 300, 260, {"depth_confidence": -1, "width_confidence": -1}, (480, 640)),
        ("sp_small_adaptive", 2, 512, 400, {}, (768, 1024)),
        ("sp_prune", 3, 1800, 1700, {"prune_min_kpts": 1536}, (1024, 1024)),
        ("din128_fixed", 4, 200, 333, {"input_dim": 128, "depth_confidence": -1, "width_confidence": -1}, (600, 800)),
        ("tiny", 5, 9, 17, {"depth_confidence": -1, "width_confidence": -1}, (100, 120)),
        ("cfg2_2048_adaptive", 6, 2048, 2048, {}, (1024, 1024)),
        ("prune_only", 7, 1700, 1650, {"depth_confidence": -1}, (1024, 1024)),
    ]
    for name, seed, m, n, over, size_hw in cases:
        conf = {**o_lg.DEFAULT_CONF, **over}
        w = o_lg.seeded_weights(conf, seed=seed)
        f0, f1 = lg_pair(seed, m, n, conf["input_dim"], size_hw)
        ref = run_ref_lg(lgmod, w, conf, f0, f1)
        ora = o_lg.match(f0, f1, w, conf)
        assert ref["stop"] == ora["stop"], (name, ref["stop"], ora["stop"])
        assert np.array_equal(ref["matches"], ora["matches"]), name
        ds = np.abs(ref["scores"] - ora["scores"]).max() if len(ref["scores"]) else 0.0
        assert np.array_equal(ref["prune0"], ora["prune0"]) and np.array_equal(ref["prune1"], ora["prune1"]), name
        print(f"  [{name}] matches={len(ref['matches'])} stop={ref['stop']} max|dscore|={ds:.2e} "
              f"final n0={ora.get('n_final0')} n1={ora.get('n_final1')}")
        assert ds < 5e-5  # fp32 evaluation-order noise (einsum vs matmul) through the 13x final_proj
        out[name + ".args"] = np.array([seed, m, n, size_hw[0], size_hw[1]], np.int64)
        out[name + ".conf"] = np.array([conf["input_dim"], conf["depth_confidence"], conf["width_confidence"],
                                        conf["prune_min_kpts"]], np.float64)
        out[name + ".matches"] = ref["matches"].astype(np.int32)
        out[name + ".scores"] = ref["scores"].astype(np.float32)
        out[name + ".stop"] = np.array(ref["stop"])
        out[name + ".prune0"] = ref["prune0"].astype(np.int8)
        out[name + ".prune1"] = ref["prune1"].astype(np.int8)
    np.savez_compressed(os.path.join(GOLD, "lightglue_golden.npz"), **out)



def gen_nn():
    """hloc cosine mutual NN is in-tree -> pinned; kornia modes are restated only (unpinned)."""
    import types
    pkg = types.ModuleType("hlocpkg"); pkg.__path__ = [T + "hloc"]
    utils = types.ModuleType("hlocpkg.utils"); utils.__path__ = [T + "hloc/utils"]
    sys.modules.update({"hlocpkg": pkg, "hlocpkg.utils": utils})
    load_by_path("hlocpkg.utils.base_model", T + "hloc/utils/base_model.py")
    matchers = types.ModuleType("hlocpkg.matchers"); matchers.__path__ = [T + "hloc/matchers"]
    sys.modules["hlocpkg.matchers"] = matchers
    nnmod = load_by_path("hlocpkg.matchers.nearest_neighbor", T + "hloc/matchers/nearest_neighbor.py")
    rng = np.random.default_rng(11)
    out = {}
    for name, n, m, ratio in [("plain", 700, 650, None), ("ratio", 512, 777, 0.9)]:
        a = rng.standard_normal((128, n)).astype(np.float32); a /= np.linalg.norm(a, axis=0)
        b = rng.standard_normal((128, m)).astype(np.float32); b /= np.linalg.norm(b, axis=0)
        k = min(n, m) // 2
        b[:, :k] = a[:, rng.permutation(n)[:k]] + 0.3 * rng.standard_normal((128, k)).astype(np.float32)
        b /= np.linalg.norm(b, axis=0)
        a = a.astype(np.float16).astype(np.float32); b = b.astype(np.float16).astype(np.float32)
        net = nnmod.NearestNeighbor({"ratio_threshold": ratio, "distance_threshold": None, "do_mutual_check": True})
        with torch.no_grad():
            ref = net({"descriptors0": torch.from_numpy(a)[None], "descriptors1": torch.from_numpy(b)[None]})
        m0 = ref["matches0"][0].numpy()
        om0, _ = o_nn.hloc_mutual_nn(a, b, ratio_thresh=ratio)
        assert np.array_equal(m0, om0), name
        print(f"  [hloc_nn {name}] matched={int((m0 > -1).sum())}/{n}")
        out[name + ".args"] = np.array([n, m, -1 if ratio is None else ratio], np.float64)
        out[name + ".desc0"] = a.astype(np.float16)
        out[name + ".desc1"] = b.astype(np.float16)
        out[name + ".matches0"] = m0.astype(np.int32)
    np.savez_compressed(os.path.join(GOLD, "nn_golden.npz"), **out)


def gen_aliked():
    """ALIKED-n16rot: the reference's LightGlue port (kornia / .utils stubbed as in SURVEY Appendix E) vs the oracle."""
    import types
    from oracle import aliked as o_al
    sd = torch.load(T + "ALIKED/models/aliked-n16rot.pth", map_location="cpu")
    w = {k: v.numpy().astype(np.float32) for k, v in sd.items() if v.dtype.is_floating_point}
    np.savez(os.path.join(DATA, "aliked_n16rot_weights.npz"), **w)
    sd16 = torch.load(T + "ALIKED/models/aliked-n16.pth", map_location="cpu")  # same architecture, the model default (aliked.py:563)
    np.savez(os.path.join(DATA, "aliked_n16_weights.npz"), **{k: v.numpy().astype(np.float32) for k, v in sd16.items() if v.dtype.is_floating_point})
    k = types.ModuleType("kornia"); kc = types.ModuleType("kornia.color")
    kc.grayscale_to_rgb = lambda x: x.repeat(1, 3, 1, 1)
    k.color = kc
    sys.modules.update({"kornia": k, "kornia.color": kc})
    pkg = types.ModuleType("lgpkg"); pkg.__path__ = [T + "LightGlue/lightglue"]
    u = types.ModuleType("lgpkg.utils")

    class Extractor(torch.nn.Module):  # mirrors thirdparty/LightGlue/lightglue/utils.py:128-131
        def __init__(self, **conf):
            super().__init__()
            self.conf = types.SimpleNamespace(**{**self._default_conf, **conf})
    u.Extractor = Extractor
    sys.modules.update({"lgpkg": pkg, "lgpkg.utils": u})
    torch.hub.load_state_dict_from_url = lambda *a, **kw: sd
    al = load_by_path("lgpkg.aliked", T + "LightGlue/lightglue/aliked.py")
    rgb = cv2.cvtColor(cv2.imread("/root/reference/assets/pytest/images/DSC_6466.jpg"), cv2.COLOR_BGR2RGB)
    cases = {
        "real224x288": (rgb[100:324, 200:488].astype(np.float32), {"max_num_keypoints": 4000, "detection_threshold": 0.2, "nms_radius": 2}),
        "real_odd203x260_r3_top100": (rgb[60:263, 100:360].astype(np.float32), {"max_num_keypoints": 100, "detection_threshold": 0.2, "nms_radius": 3}),
        "blocks256": (synthetic.blocks_image(5, 256).astype(np.float32), {"max_num_keypoints": 4096, "detection_threshold": 0.2, "nms_radius": 3}),
    }
    out = {}
    for name, (img, over) in cases.items():
        conf = {**o_al.DEFAULT_CONF, **over}
        net = al.ALIKED(**conf).eval()
        with torch.no_grad():
            x = torch.tensor(img.transpose(2, 0, 1)[None] / 255.0, dtype=torch.float)
            r = net({"image": x})
        ref = {"keypoints": r["keypoints"][0].numpy(), "descriptors": r["descriptors"][0].numpy().T, "scores": r["keypoint_scores"][0].numpy()}
        ora = o_al.extract(img, w, conf)
        assert ref["keypoints"].shape == ora["keypoints"].shape, (name, ref["keypoints"].shape, ora["keypoints"].shape)
        dk = np.abs(ref["keypoints"] - ora["keypoints"]).max()
        dd = np.abs(ref["descriptors"] - ora["descriptors"]).max()
        ds = np.abs(ref["scores"] - ora["scores"]).max()
        print(f"  [aliked {name}] N={len(ref['keypoints'])} max|dkpt|={dk:.2e} max|ddesc|={dd:.2e} max|dscore|={ds:.2e} score range {ref['scores'].min():.3f}..{ref['scores'].max():.3f}")
        assert dk < 1e-4 and dd < 1e-5 and ds < 1e-5
        out[name + ".image"] = img.astype(np.uint8)
        out[name + ".conf"] = np.array([conf["max_num_keypoints"], conf["detection_threshold"], conf["nms_radius"]], np.float64)
        out[name + ".keypoints"] = ref["keypoints"].astype(np.float32)
        out[name + ".scores"] = ref["scores"].astype(np.float32)
        out[name + ".descriptors"] = ref["descriptors"].astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "aliked_golden.npz"), **out)


def gen_lighterglue():
    """Trained-weights known-answer test of the LightGlue algorithm: XFeat features (reference module,
    weights/xfeat.pt) of the reference's own test photos DSC_6466 / DSC_6467, matched by the vendored LightGlue class
    carrying the trained LighterGlue checkpoint (weights/xfeat-lighterglue.pt: descriptor_dim 96, one head, 6 layers,
    input_dim 64; key renames of modules/lighterglue.py:40-46), against the oracle."""
    XF = T + "accelerated_features"
    sys.path.insert(0, XF)
    from modules.xfeat import XFeat  # noqa: E402  (torch-only module of the reference)
    xf = XFeat(weights=XF + "/weights/xfeat.pt", top_k=2048)
    xf.dev = torch.device("cpu"); xf.net.to("cpu")
    feats = []
    for name in ("DSC_6466.jpg", "DSC_6467.jpg"):
        bgr = cv2.imread("/root/reference/assets/pytest/images/" + name)
        x = torch.from_numpy(bgr[..., ::-1].copy()).permute(2, 0, 1)[None].float()
        o = xf.detectAndCompute(x, top_k=2048)[0]
        k = o["keypoints"].numpy().astype(np.float32)
        d = o["descriptors"].numpy().astype(np.float32)
        # FeaturesDict after the fp16 h5 round trip, (D,N) layout, image_size [H,W] as DIM passes it (quirk A.3)
        feats.append({"keypoints": k.astype(np.float16).astype(np.float32), "descriptors": d.T.astype(np.float16).astype(np.float32).copy(),
                      "image_size": np.array(bgr.shape[:2], np.int32)})
        print(f"  [xfeat {name}] {len(k)} kpts, desc {d.shape}")
    sd = torch.load(XF + "/weights/xfeat-lighterglue.pt", map_location="cpu")
    sd = {k: v for k, v in sd.items() if k.startswith("matcher.")}  # the checkpoint also carries the XFeat extractor
    for i in range(6):
        sd = {k.replace(f"self_attn.{i}", f"transformers.{i}.self_attn"): v for k, v in sd.items()}
        sd = {k.replace(f"cross_attn.{i}", f"transformers.{i}.cross_attn"): v for k, v in sd.items()}
    sd = {k.replace("matcher.", ""): v for k, v in sd.items()}
    w = {k: v.numpy().astype(np.float32) for k, v in sd.items() if v.dtype.is_floating_point and k != "confidence_thresholds"}
    np.savez_compressed(os.path.join(DATA, "lighterglue_weights.npz"), **w)
    lgmod = load_by_path("ref_lightglue", T + "LightGlue/lightglue/lightglue.py")
    out = {"kpts0": feats[0]["keypoints"].astype(np.float16), "kpts1": feats[1]["keypoints"].astype(np.float16),
           "desc0": feats[0]["descriptors"].astype(np.float16), "desc1": feats[1]["descriptors"].astype(np.float16),
           "size0": feats[0]["image_size"], "size1": feats[1]["image_size"]}
    base = {**o_lg.DEFAULT_CONF, "input_dim": 64, "descriptor_dim": 96, "num_heads": 1, "n_layers": 6}
    for name, over in [("fixed", {"depth_confidence": -1, "width_confidence": -1}),
                       ("lighterglue_default", {"depth_confidence": -1, "width_confidence": 0.95}),   # modules/lighterglue.py:12-27
                       ("dim_plugin_default", {"depth_confidence": 0.95, "width_confidence": 0.99})]:  # matchers/lighterglue.py:79-86
        conf = {**base, **over}
        ref = run_ref_lg(lgmod, w, conf, feats[0], feats[1])
        ora = o_lg.match(feats[0], feats[1], w, conf)
        ds = np.abs(ref["scores"] - ora["scores"]).max() if len(ref["scores"]) else 0.0
        print(f"  [lighterglue {name}] matches={len(ref['matches'])} stop={ref['stop']} mean score={ref['scores'].mean():.3f} "
              f"oracle: matches={len(ora['matches'])} stop={ora['stop']} max|dscore|={ds:.2e}")
        assert ref["stop"] == ora["stop"] and np.array_equal(ref["matches"], ora["matches"]), name
        assert np.array_equal(ref["prune0"], ora["prune0"]) and np.array_equal(ref["prune1"], ora["prune1"]), name
        assert ds < 2e-4  # fp32 evaluation-order noise (einsum vs matmul) with trained weights, scores up to 1
        out[name + ".conf"] = np.array([conf["depth_confidence"], conf["width_confidence"]], np.float64)
        out[name + ".matches"] = ref["matches"].astype(np.int32)
        out[name + ".scores"] = ref["scores"].astype(np.float32)
        out[name + ".stop"] = np.array(ref["stop"])
        out[name + ".prune0"] = ref["prune0"].astype(np.int8)
        out[name + ".prune1"] = ref["prune1"].astype(np.int8)
    np.savez_compressed(os.path.join(GOLD, "lighterglue_golden.npz"), **out)


def run_ref_sg(net, f0, f1):
    """Drive the reference SuperGlue class exactly as SuperGlueMatcher._match_pairs does (features_2_sg, superglue.py:8-41)."""
    data = {}
    for i, f in enumerate((f0, f1)):
        for k in ("keypoints", "descriptors", "scores"):
            data[f"{k}{i}"] = torch.tensor(np.asarray(f[k]), dtype=torch.float)[None]
        h, w = [int(v) for v in f["image_size"]]
        data[f"image{i}"] = torch.empty(1, 1, h, w)
    with torch.no_grad():
        r = net(data)
    return {"matches0": r["matches0"][0].numpy().astype(np.int64), "matching_scores0": r["matching_scores0"][0].numpy()}


def gen_superglue():
    """SuperGlue: the oracle against the reference class with (a) its vendored TRAINED outdoor weights on
    SuperPoint features of two overlapping crops of the reference's test photo (checked here, not stored: the checkpoint is
    48 MB) and (b) seeded weights on seeded features (stored as golden vectors)."""
    from oracle import superglue as o_sg
    sgmod = load_by_path("ref_superglue", T + "SuperGluePretrainedNetwork/models/superglue.py")
    net = sgmod.SuperGlue({"weights": "outdoor"}).eval()
    wt = {k: v.numpy().astype(np.float32) for k, v in net.state_dict().items() if v.dtype.is_floating_point}
    _, spw = sp_weights()
    rgb = cv2.cvtColor(cv2.imread("/root/reference/assets/pytest/images/DSC_6466.jpg"), cv2.COLOR_BGR2RGB)
    gray = synthetic.to_gray_like_reference(rgb)
    conf = {"nms_radius": 3, "keypoint_threshold": 0.0005, "max_keypoints": 512}
    feats = []
    for crop in (gray[100:340, 200:520], gray[120:360, 230:550]):
        f = o_sp.extract(crop.copy(), spw, conf)
        f["image_size"] = np.array(crop.shape[:2], np.int32)
        feats.append(f)
    ref = run_ref_sg(net, feats[0], feats[1])
    ora = o_sg.match(feats[0], feats[1], wt)
    nm = int((ref["matches0"] > -1).sum())
    ds = np.abs(ref["matching_scores0"] - ora["matching_scores0"]).max()
    print(f"  [superglue trained outdoor, real crops] matches={nm} oracle matches={len(ora['matches'])} max|dscore|={ds:.2e}")
    assert np.array_equal(ref["matches0"], ora["matches0"]) and nm > 100 and ds < 2e-4
    out = {}
    for name, seed, m, n, size_hw in [("small", 1, 300, 260, (480, 640)), ("mid", 2, 512, 400, (768, 1024)), ("tiny", 3, 9, 17, (100, 120))]:
        w = o_sg.seeded_weights(seed)
        missing = net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
        assert not missing.unexpected_keys and all("num_batches_tracked" in k for k in missing.missing_keys), missing
        f0, f1 = lg_pair(seed, m, n, 256, size_hw)
        ref = run_ref_sg(net, f0, f1)
        ora = o_sg.match(f0, f1, w)
        nm = int((ref["matches0"] > -1).sum())
        ds = np.abs(ref["matching_scores0"] - ora["matching_scores0"]).max()
        print(f"  [superglue seeded {name}] matches={nm} max|dscore|={ds:.2e}")
        assert np.array_equal(ref["matches0"], ora["matches0"]) and ds < 1e-4
        out[name + ".args"] = np.array([seed, m, n, size_hw[0], size_hw[1]], np.int64)
        out[name + ".matches0"] = ref["matches0"].astype(np.int32)
        out[name + ".matching_scores0"] = ref["matching_scores0"].astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "superglue_golden.npz"), **out)


def gen_cfg1():
    """BASELINE config 1 (reference CPU plumbing): sift + kornia_matcher on assets/example_sacre_coeur A / B, exactly as
    ExtractorBase.extract -> SIFTExtractor._extract feeds it (extractor_base.py:190-202 gray quirk, as_float False;
    extractors/sift.py:27-50 with config.py:234-242).  Stored: the SIFT features (what DIM writes to features.h5 before the
    fp16 cast) and the smnn-0.85 matches of the oracle on their fp16 round trip (kornia itself is not installable: unpinned)."""
    sift = cv2.SIFT_create(nfeatures=2048, nOctaveLayers=3, contrastThreshold=0.0004, edgeThreshold=10, sigma=1.6)
    out, feats = {}, []
    for tag, name in (("0", "sacre_coeur_A.jpg"), ("1", "sacre_coeur_B.jpg")):
        rgb = cv2.cvtColor(cv2.imread("/root/reference/assets/example_sacre_coeur/images/" + name), cv2.COLOR_BGR2RGB)
        gray = cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY)  # sic; as_float False -> uint8 goes to SIFT
        kp, des = sift.detectAndCompute(gray, None)
        kpts = cv2.KeyPoint_convert(kp).astype(np.float32)
        des = des.astype(float).T
        assert np.array_equal(des, np.round(des)) and des.max() <= 255
        out["kpts" + tag], out["desc" + tag], out["size" + tag] = kpts, des.astype(np.uint8), np.array(gray.shape[:2], np.int32)
        half = lambda a: a.astype(np.float16).astype(np.float32)
        feats.append({"keypoints": half(kpts), "descriptors": half(des.astype(np.float32))})
        print(f"  [cfg1 {name}] {gray.shape} -> {len(kpts)} SIFT keypoints")
    idx, dist = o_nn.kornia_match(feats[0], feats[1], "smnn", 0.85)
    print(f"  [cfg1] smnn 0.85: {len(idx)} matches")
    out["smnn085.matches"], out["smnn085.dist"] = idx.astype(np.int32), dist.astype(np.float32)
    idx, dist = o_nn.kornia_match(feats[0], feats[1], "mnn")
    out["mnn.matches"] = idx.astype(np.int32)
    np.savez_compressed(os.path.join(GOLD, "cfg1_sift_golden.npz"), **out)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["sp", "lg", "nn", "aliked", "lighterglue", "superglue", "cfg1"]
    if "sp" in which:
        print("SuperPoint: reference vs oracle"); gen_superpoint()
    if "lg" in which:
        print("LightGlue: reference vs oracle"); gen_lightglue()
    if "nn" in which:
        print("NN: reference(hloc) vs oracle"); gen_nn()
    if "aliked" in which:
        print("ALIKED: reference vs oracle"); gen_aliked()
    if "lighterglue" in which:
        print("LighterGlue (trained weights): reference vs oracle"); gen_lighterglue()
    if "superglue" in which:
        print("SuperGlue: reference vs oracle"); gen_superglue()
    if "cfg1" in which:
        print("cfg1: OpenCV SIFT on sacre_coeur A/B + oracle smnn"); gen_cfg1()
    print("golden fixtures written to", GOLD)
