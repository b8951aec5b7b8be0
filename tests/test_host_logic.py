"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol, the plugin
surface mirrors the reference's, weight packing, the features.h5 boundary, loud failure without CUDA."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from dim_b200 import _native
    lib = _native.load_library()
    header = open(os.path.join(ROOT, "include", "dimb200.h")).read()
    declared = set(re.findall(r"\b(dimb_[a-z0-9_]+)\s*\(", header))
    assert declared >= set(_native.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert b"sm_100a" in lib.dimb_version()


def test_no_cpu_fallback():
    import torch
    from dim_b200 import _native
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.DimbError, match="no CPU fallback"):
        _native.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "deep-image-matching_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_plugin_surface():
    import dim_b200.extractors as E
    import dim_b200.matchers as M
    from dim_b200.config import Config
    from dim_b200.extractors.extractor_base import ExtractorBase, extractor_loader
    from dim_b200.matchers.matcher_base import MatcherBase, matcher_loader
    sp = extractor_loader(E, "superpoint")
    lg = matcher_loader(M, "lightglue")
    km = matcher_loader(M, "kornia_matcher")
    ltg = matcher_loader(M, "lighterglue")
    assert ltg.__name__ == "LighterGlueMatcher" and issubclass(ltg, MatcherBase) and ltg.min_matches == 20
    sg = matcher_loader(M, "superglue")
    assert sg.__name__ == "SuperGlueMatcher" and sg.max_feat_no_tiling == 50000 and sg.default_config["sinkhorn_iterations"] == 20
    al = extractor_loader(E, "aliked")
    assert al.__name__ == "AlikedExtractor" and issubclass(al, ExtractorBase)
    assert al.grayscale is False and al.descriptor_size == 128 and al._default_conf["nms_radius"] == 2
    assert sp.__name__ == "SuperPointExtractor" and issubclass(sp, ExtractorBase)
    assert lg.__name__ == "LightGlueMatcher" and issubclass(lg, MatcherBase)
    assert km.__name__ == "KorniaMatcher"
    assert sp.grayscale and sp.descriptor_size == 256 and sp.features_as_half
    assert lg.max_feat_no_tiling == 200000 and lg.min_matches == 20
    assert sp._default_conf["fix_sampling"] is False and sp._default_conf["nms_radius"] == 4
    with pytest.raises(TypeError, match="Config object"):
        sp({"extractor": {}})
    with pytest.raises(TypeError, match="Config object"):
        lg({"matcher": {}})
    cfg = Config(pipeline="superpoint+lightglue")
    assert cfg.extractor["max_keypoints"] == 2048 and cfg.matcher["depth_confidence"] == 0.95


def test_featuresdict_layout_decision():
    from dim_b200.matchers.lightglue import featuresDict2Lightglue
    k = np.zeros((10, 2), np.float32)
    assert featuresDict2Lightglue({"keypoints": k, "descriptors": np.zeros((256, 10), np.float32)})["_layout"] == 0
    assert featuresDict2Lightglue({"keypoints": k, "descriptors": np.zeros((10, 256), np.float32)})["_layout"] == 1
    with pytest.raises(ValueError, match="mismatch"):
        featuresDict2Lightglue({"keypoints": k, "descriptors": np.zeros((7, 256), np.float32)})
    with pytest.raises(KeyError):
        featuresDict2Lightglue({"keypoints": k})
    # N == D is ambiguous in the reference too: treated as (N,D)   
    k2 = np.zeros((256, 2), np.float32)
    assert featuresDict2Lightglue({"keypoints": k2, "descriptors": np.zeros((256, 256), np.float32)})["_layout"] == 1


def test_weight_packing_sizes(sp_weights):
    from dim_b200 import _native, weights
    assert _native.pack_superpoint_weights(sp_weights).size == 1300865
    w = weights.lightglue_seeded()
    assert _native.pack_lightglue_weights(w, 256, 256, 9).size == 11851601
    w128 = weights.lightglue_seeded(input_dim=128)
    assert _native.pack_lightglue_weights(w128, 128, 256, 9).size == 11851601 + 256 * 128 + 256
    # old-style checkpoint prefixes are renamed (lightglue.py:391-396)
    old = {k.replace("transformers.3.self_attn", "self_attn.3").replace("transformers.3.cross_attn", "cross_attn.3"): v for k, v in w.items()}
    assert np.array_equal(_native.pack_lightglue_weights(old, 256, 256, 9), _native.pack_lightglue_weights(w, 256, 256, 9))


def test_aliked_and_lighterglue_weight_packing(al_weights, ltg_weights):
    """Blob sizes the C ABI checks: ALIKED-n16 state_dict order (678316 floats), LighterGlue shape (64 / 96 / 1 head / 6 layers)."""
    from dim_b200 import _native
    assert _native.pack_aliked_weights(al_weights).size == 678316
    assert len(_native.aliked_weight_names()) == 68 and set(_native.aliked_weight_names()) == set(al_weights)
    d, din, L = 96, 64, 6
    per_layer = (3 * d * d + 3 * d) + (d * d + d) + (4 * d * d + 2 * d) + 4 * d + (2 * d * d + d) + 3 * (d * d + d) + (4 * d * d + 2 * d) + 4 * d + (2 * d * d + d)
    need = (d // 2) * 2 + d * din + d + per_layer * L + L * (d + 1 + d * d + d) + (L - 1) * (d + 1)
    assert _native.pack_lightglue_weights(ltg_weights, din, d, L).size == need


def test_seeded_weights_are_deterministic():
    from dim_b200 import weights
    a, b = weights.lightglue_seeded(seed=3), weights.lightglue_seeded(seed=3)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert abs(float(a["transformers.0.self_attn.Wqkv.weight"][0, 0]) - float(weights.lightglue_seeded(seed=4)["transformers.0.self_attn.Wqkv.weight"][0, 0])) > 0


def test_h5_boundary_roundtrip(tmp_path):
    from dim_b200.io_h5 import as_half_roundtrip, get_features, save_features_h5
    rng = np.random.default_rng(0)
    feats = {"keypoints": rng.uniform(0, 3000, (50, 2)).astype(np.float32), "descriptors": rng.standard_normal((256, 50)).astype(np.float32),
             "scores": rng.uniform(0, 1, 50).astype(np.float32), "tile_idx": np.zeros(50, np.float32), "image_size": np.array([1536, 2048])}
    save_features_h5(tmp_path / "features.h5", dict(feats), "a.jpg")
    back = get_features(tmp_path / "features.h5", "a.jpg")
    exp = as_half_roundtrip(feats)
    for k in feats:
        assert np.array_equal(back[k], exp[k]), k
    assert back["image_size"].dtype == np.int32 and back["keypoints"].dtype == np.float32
    assert np.abs(back["keypoints"] - feats["keypoints"]).max() <= 1.0  # fp16 quantisation above 2048 px
    with pytest.raises(ValueError):
        get_features(tmp_path / "features.h5", "missing.jpg")


def test_synthetic_generator_is_deterministic():
    from dim_b200 import synthetic
    a0, a1 = synthetic.synthetic_pair(5, 256)
    b0, b1 = synthetic.synthetic_pair(5, 256)
    assert np.array_equal(a0, b0) and np.array_equal(a1, b1) and a0.dtype == np.float32 and a0.shape == (256, 256)


def test_pair_generators():
    """pairs_from_sequential / pairs_from_bruteforce restate pairs_generator.py:22-38."""
    from dim_b200.pairs_generator import pairs_from_bruteforce, pairs_from_sequential
    imgs = [f"im{i}.jpg" for i in range(5)]
    assert pairs_from_sequential(imgs, 1) == [(imgs[i], imgs[i + 1]) for i in range(4)]
    assert pairs_from_sequential(imgs, 2) == [("im0.jpg", "im1.jpg"), ("im0.jpg", "im2.jpg"), ("im1.jpg", "im2.jpg"), ("im1.jpg", "im3.jpg"),
                                              ("im2.jpg", "im3.jpg"), ("im2.jpg", "im4.jpg"), ("im3.jpg", "im4.jpg")]
    assert len(pairs_from_bruteforce(imgs)) == 10 and pairs_from_bruteforce(imgs)[0] == ("im0.jpg", "im1.jpg")
    assert len(pairs_from_sequential([f"d{i}" for i in range(200)], 1)) == 199  # cfg5


def test_c_abi_rejects_null_handles_without_touching_the_gpu():
    """Argument validation comes before any CUDA call: NULL handles / buffers return DIMB_ERR_ARG (-3) on a machine without a GPU."""
    import ctypes as C
    from dim_b200 import _native
    lib = _native.load_library()
    null, n = C.c_void_p(), C.c_int(0)
    buf = np.zeros(16, np.float32)
    assert lib.dimb_sp_create(null, null, 0, None, C.byref(C.c_void_p())) == -3
    assert lib.dimb_lg_create(null, null, 0, None, C.byref(C.c_void_p())) == -3
    assert lib.dimb_aliked_create(null, null, 0, None, C.byref(C.c_void_p())) == -3
    assert lib.dimb_sg_create(null, null, 0, None, C.byref(C.c_void_p())) == -3
    assert lib.dimb_sp_extract(null, _native._ptr(buf), 1, 16, 16, null, null, null, null, 1) == -3
    assert lib.dimb_aliked_extract(null, _native._ptr(buf), 16, 16, 3, null, null, null, null, 1) == -3
    assert lib.dimb_lg_match(null, 1, None, None, null, null, null, null, 1) == -3
    assert lib.dimb_nn_match(null, null, 0, null, 0, 256, 0, C.c_float(0.0), null, null, C.byref(n), 1) == -3
    assert lib.dimb_last_error(null) == b"null context"
    lib.dimb_sp_destroy(null), lib.dimb_lg_destroy(null), lib.dimb_aliked_destroy(null), lib.dimb_sg_destroy(null), lib.dimb_pipe_destroy(null)
