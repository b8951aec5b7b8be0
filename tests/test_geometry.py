"""Geometric verification: fundamental-matrix RANSAC.  The arithmetic (csrc/gv_math.cuh) is shared by the CUDA
kernels and a host driver in the self-test library, so the estimator itself is tested WITHOUT a GPU; the GPU tests check that the
kernels drive it identically and against OpenCV.  RANSAC is stochastic in the reference as well (pydegensac / OpenCV RNG): parity
is statistical - inlier sets on data with known geometry."""
import ctypes as C

import numpy as np
import pytest


def two_view(seed, n=600, outlier_every=3, noise=0.3):
    """Random 3-D points seen by two cameras; every `outlier_every`-th correspondence is replaced by a random point."""
    rng = np.random.default_rng(seed)
    f, c = 800.0, np.array([512.0, 384.0])
    X = np.stack([3 * rng.uniform(-1, 1, n), 2 * rng.uniform(-1, 1, n), 6 + 2 * rng.uniform(-1, 1, n)], 1)
    a = 0.09
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    t = np.array([0.6, 0.05, 0.1])
    k0 = f * X[:, :2] / X[:, 2:] + c
    Xc = X @ R.T + t
    k1 = f * Xc[:, :2] / Xc[:, 2:] + c + noise * rng.uniform(-1, 1, (n, 2))
    gt = np.ones(n, bool)
    out = np.arange(n) % outlier_every == 0
    k1[out] = c + rng.uniform(-1, 1, (int(out.sum()), 2)) * np.array([500, 380])
    gt[out] = False
    return k0.astype(np.float32), k1.astype(np.float32), gt


def check_model(F, mask, k0, k1, gt):
    assert F is not None and mask.dtype == bool and mask.shape == gt.shape
    assert (mask & gt).sum() >= 0.98 * gt.sum()             # recovers the true inliers
    assert (mask & ~gt).sum() <= 0.04 * (~gt).sum() + 2     # an outlier passes only by lying next to its epipolar line
    h0 = np.concatenate([k0, np.ones((len(k0), 1), np.float32)], 1).astype(np.float64)
    h1 = np.concatenate([k1, np.ones((len(k1), 1), np.float32)], 1).astype(np.float64)
    Fd = np.asarray(F, np.float64)
    e = np.einsum("ni,ij,nj->n", h1, Fd, h0)
    l0, l1 = h0 @ Fd.T, h1 @ Fd
    samp = e ** 2 / (l0[:, 0] ** 2 + l0[:, 1] ** 2 + l1[:, 0] ** 2 + l1[:, 1] ** 2)
    assert np.all(samp[mask] < 1.0 + 1e-3) and abs(np.linalg.det(Fd)) < 1e-6  # x1^T F x0 = 0 convention, rank 2


def test_ransac_arithmetic_on_the_host():
    """csrc/gv_math.cuh driven on the CPU by the self-test library (no GPU): known two-view geometry with 33 % outliers."""
    from dim_b200 import _native
    lib = _native.load_selftest_library()
    for seed in (1, 2, 3):
        k0, k1, gt = two_view(seed)
        F, mask = np.zeros(9, np.float32), np.zeros(len(k0), np.uint8)
        rc = lib.dimb_gv_host(k0.ctypes.data, k1.ctypes.data, len(k0), C.c_float(1.0), 1500, seed, F.ctypes.data, mask.ctypes.data)
        assert rc == 0
        check_model(F.reshape(3, 3), mask.astype(bool), k0, k1, gt)
    assert lib.dimb_gv_host(k0.ctypes.data, k1.ctypes.data, 7, C.c_float(1.0), 10, 0, F.ctypes.data, mask.ctypes.data) == -3


def test_reference_signature_without_gpu():
    """Contract of geometric_verification (utils/geometric_verification.py:45-111) that needs no estimator."""
    from dim_b200.geometric_verification import geometric_verification
    k = np.zeros((5, 2), np.float32)
    F, m = geometric_verification(k, k, method="NONE")
    assert F is None and m.all() and m.dtype == bool
    F, m = geometric_verification(k, k, method="pydegensac")  # fewer than 8 matches: nothing to verify
    assert F is None and m.shape == (5,) and m.all()
    with pytest.raises(ValueError, match="Invalid Geometry Verification method"):
        geometric_verification(k, k, method="bogus")


@pytest.mark.gpu
def test_gpu_ransac_known_geometry_and_opencv(ctx):
    import cv2
    from dim_b200.geometric_verification import geometric_verification
    for seed in (1, 2, 3, 4):
        k0, k1, gt = two_view(seed, n=1500)
        F, mask = geometric_verification(k0, k1, method="pydegensac", threshold=1.0, max_iters=4096, seed=seed)
        check_model(F, mask, k0, k1, gt)
        _, inl = cv2.findFundamentalMat(k0, k1, cv2.RANSAC, 1.0, 0.9999, 10000)
        cvm = inl.ravel() > 0
        assert (mask & cvm).sum() >= 0.97 * cvm.sum()  # the same correspondences as OpenCV's RANSAC (statistical parity)
        # reproducible: same seed, same answer
        F2, mask2 = geometric_verification(k0, k1, method="ransac", threshold=1.0, max_iters=4096, seed=seed)
        assert np.array_equal(mask, mask2) and np.allclose(F, F2)
    F, mask = geometric_verification(k0[:7], k1[:7])
    assert F is None and mask.all()


@pytest.mark.gpu
def test_gpu_ransac_batch_on_match_tables(ctx):
    """dimb_gv_fundamental_batch_dev on the device layouts the matchers produce ([P][cap][2] int64 matches + counts indexing the
    keypoint arrays): per pair equal to the single-pair host entry with the same seed offset."""
    import torch
    from dim_b200 import _native
    P, cap = 3, 1024
    kp0, kp1, m, nm, gts = [], [], torch.zeros(P, cap, 2, dtype=torch.int64, device="cuda"), torch.zeros(P, dtype=torch.int32, device="cuda"), []
    for p in range(P):
        k0, k1, gt = two_view(10 + p, n=700 + 100 * p)
        perm = np.random.default_rng(p).permutation(len(k0))  # matches index shuffled keypoint arrays
        kp0.append(torch.from_numpy(k0).cuda())
        kp1.append(torch.from_numpy(k1[perm]).cuda())
        inv = np.argsort(perm)
        m[p, :len(k0), 0] = torch.arange(len(k0))
        m[p, :len(k0), 1] = torch.from_numpy(inv)
        nm[p] = len(k0)
        gts.append((k0, k1, gt))
    F = torch.zeros(P, 9, device="cuda"); mask = torch.zeros(P, cap, dtype=torch.uint8, device="cuda"); ninl = torch.zeros(P, dtype=torch.int32, device="cuda")
    a0 = (C.c_void_p * P)(*[t.data_ptr() for t in kp0]); a1 = (C.c_void_p * P)(*[t.data_ptr() for t in kp1])
    ctx.check(ctx.lib.dimb_gv_fundamental_batch_dev(ctx.h, P, a0, a1, m.data_ptr(), nm.data_ptr(), cap, 1.0, 4096, 5, F.data_ptr(), mask.data_ptr(),
                                                    ninl.data_ptr(), torch.cuda.current_stream().cuda_stream), "dimb_gv_fundamental_batch_dev")
    torch.cuda.synchronize()
    for p, (k0, k1, gt) in enumerate(gts):
        mk = mask[p, :len(k0)].cpu().numpy().astype(bool)
        assert int(ninl[p]) == mk.sum()
        check_model(F[p].cpu().numpy().reshape(3, 3), mk, k0, k1, gt)
