"""The CPU oracle against the golden vectors produced by the reference's own model code
(oracle/gen_golden.py, run in the authoring container where /root/reference exists)."""
import numpy as np
import pytest

from conftest import AL_CASES, LG_CASES, LTG_CASES, SP_CASES, al_case, lg_case, ltg_case, sp_case
from oracle import lightglue as o_lg
from oracle import nn_match as o_nn
from oracle import superpoint as o_sp


@pytest.mark.parametrize("name", SP_CASES)
def test_superpoint_oracle_matches_reference(name, sp_golden, sp_weights):
    img, conf, ref = sp_case(sp_golden, name)
    out = o_sp.extract(img, sp_weights, conf)
    a, b = o_sp.canonical_order(out), o_sp.canonical_order(ref)
    assert np.array_equal(out["keypoints"][a], ref["keypoints"][b])
    assert np.abs(out["scores"][a] - ref["scores"][b]).max() < 2e-6
    assert np.abs(out["descriptors"][:, a] - ref["descriptors"][:, b]).max() < 2e-6


def test_superpoint_oracle_cfg2_full_size(sp_golden, sp_weights):
    from dim_b200 import synthetic
    g0, _ = synthetic.synthetic_pair(0, 1024)
    out = o_sp.extract(g0, sp_weights, {"nms_radius": 3, "keypoint_threshold": 0.0005, "max_keypoints": 2048})
    a = o_sp.canonical_order(out)
    assert len(a) == 2048
    assert np.array_equal(out["keypoints"][a].astype(np.int16), sp_golden["cfg2.keypoints"])
    assert np.abs(out["scores"][a] - sp_golden["cfg2.scores"]).max() < 2e-6
    assert np.abs(out["descriptors"][:, a[:64]] - sp_golden["cfg2.descriptors_first64"]).max() < 2e-6


@pytest.mark.parametrize("name", AL_CASES)
def test_aliked_oracle_matches_reference(name, al_golden, al_weights):
    """oracle/aliked.py against the outputs of the reference's own ALIKED module (same torch build: bit-identical
    in the authoring container; 1e-5 leaves room for a different CPU's conv kernels)."""
    from oracle import aliked as o_al
    from oracle.compare import compare_aliked
    img, conf, ref = al_case(al_golden, name)
    out = o_al.extract(img, al_weights, conf)
    rep = compare_aliked(out, ref, None, conf["detection_threshold"], conf["nms_radius"], tol=1e-5, tol_kpt=1e-4)
    assert rep["n"] == len(ref["keypoints"]) and rep["order_same"]
    if name.endswith("top100"):
        assert rep["n"] == 100  # the n_limit branch fired


@pytest.mark.parametrize("name", [c for c in LG_CASES if c != "cfg2_2048_adaptive"])
def test_lightglue_oracle_matches_reference(name, lg_golden):
    f0, f1, conf, w, ref = lg_case(lg_golden, name)
    out = o_lg.match(f0, f1, w, conf)
    assert out["stop"] == ref["stop"]
    assert np.array_equal(out["matches"], ref["matches"])
    if len(ref["scores"]):
        assert np.abs(out["scores"] - ref["scores"]).max() < 5e-5
    assert np.array_equal(out["prune0"], lg_golden[name + ".prune0"])


@pytest.mark.parametrize("name", LTG_CASES)
def test_lightglue_oracle_trained_weights(name, ltg_golden, ltg_weights):
    """Known-answer test with TRAINED weights: the LighterGlue checkpoint vendored by the reference (LightGlue architecture,
    descriptor_dim 96, one head, 6 layers) on XFeat features of the reference's own test photos; expected outputs come from
    the reference's LightGlue class.  Scores: fp32 evaluation-order noise is larger with trained weights (2e-4)."""
    f0, f1, conf, ref = ltg_case(ltg_golden, name)
    out = o_lg.match(f0, f1, ltg_weights, conf)
    assert out["stop"] == ref["stop"]
    assert np.array_equal(out["matches"], ref["matches"]) and len(ref["matches"]) > 390
    assert np.abs(out["scores"] - ref["scores"]).max() < 2e-4
    assert np.array_equal(out["prune0"], ref["prune0"]) and np.array_equal(out["prune1"], ref["prune1"])


@pytest.mark.parametrize("name", ["small", "mid", "tiny"])
def test_superglue_oracle_matches_reference(name):
    """oracle/superglue.py (specification of the SuperGlue row) against outputs of the reference's SuperGlue class
    with seeded weights; gen_golden.py additionally checks it with the vendored trained outdoor checkpoint (332 identical
    matches on two crops of the reference's test photo) - that checkpoint is 48 MB and is not shipped."""
    import os
    from conftest import GOLD
    from oracle import superglue as o_sg
    from oracle.gen_golden import lg_pair
    g = np.load(os.path.join(GOLD, "superglue_golden.npz"))
    seed, m, n, h, w = [int(x) for x in g[name + ".args"]]
    f0, f1 = lg_pair(seed, m, n, 256, (h, w))
    out = o_sg.match(f0, f1, o_sg.seeded_weights(seed))
    assert np.array_equal(out["matches0"], g[name + ".matches0"])
    assert np.abs(out["matching_scores0"] - g[name + ".matching_scores0"]).max() < 1e-4
    assert (out["matches0"] > -1).sum() == len(out["matches"]) > 0


@pytest.mark.parametrize("name", ["plain", "ratio"])
def test_hloc_mutual_nn_oracle(name, nn_golden):
    n, m, ratio = nn_golden[name + ".args"]
    a, b = nn_golden[name + ".desc0"].astype(np.float32), nn_golden[name + ".desc1"].astype(np.float32)
    m0, _ = o_nn.hloc_mutual_nn(a, b, ratio_thresh=None if ratio < 0 else float(ratio))
    assert np.array_equal(m0, nn_golden[name + ".matches0"])


def test_kornia_modes_are_consistent():
    """Internal consistency of the (unpinned) kornia restatement: mnn == smnn(th=inf-like) subset relations."""
    rng = np.random.default_rng(3)
    a = rng.standard_normal((64, 150)).astype(np.float32)
    b = np.concatenate([a[:, :70] + 0.05 * rng.standard_normal((64, 70)).astype(np.float32), rng.standard_normal((64, 60)).astype(np.float32)], 1)
    f0, f1 = {"descriptors": a}, {"descriptors": b}
    nn, _ = o_nn.kornia_match(f0, f1, "nn")
    mnn, _ = o_nn.kornia_match(f0, f1, "mnn")
    snn, _ = o_nn.kornia_match(f0, f1, "snn", 0.8)
    smnn, _ = o_nn.kornia_match(f0, f1, "smnn", 0.8)
    assert nn.shape == (150, 2)
    as_set = lambda x: {tuple(r) for r in x}
    assert as_set(mnn) <= as_set(nn) and as_set(snn) <= as_set(nn)
    assert as_set(smnn) <= as_set(mnn) & as_set(snn)
    assert len(mnn) >= 60 and np.all(np.diff(smnn[:, 0]) > 0)
    # mnn is symmetric under swapping the inputs
    mnn_t, _ = o_nn.kornia_match(f1, f0, "mnn")
    assert as_set(mnn) == {(j, i) for i, j in as_set(mnn_t)}
