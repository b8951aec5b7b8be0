"""Numerical model (numpy, CPU) of the EXACT-mode tensor-core arithmetic of csrc/gemm.cuh: every operand is stored as two fp16 planes
hi = fp16(x), lo = fp16(x - hi); the product is hi*hi + hi*lo + lo*hi accumulated in fp32 (the lo*lo term is dropped).  The test pins
the error budget that the 1e-4 parity tolerance relies on, and shows what the single-plane FAST mode gives up."""
import numpy as np


def split(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def test_split_operand_is_fp32_class():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(1 << 16) * np.exp(rng.uniform(-6, 6, 1 << 16))).astype(np.float32)
    hi, lo = split(x)
    err = np.abs((hi + lo) - x)
    # 11 + 11 significand bits (fp32 has 24) wherever lo stays in fp16's normal range; below that (|x| < ~2^-3) the absolute error
    # floors at half an fp16 subnormal step, 2^-25 - negligible against O(1) activations and weights in a dot product
    assert (err <= np.maximum(np.abs(x) * 2.0 ** -21, 2.0 ** -25)).all()
    assert np.abs(lo).max() <= np.abs(hi).max() * 2.0 ** -10


def test_three_term_product_vs_fast_mode():
    rng = np.random.default_rng(1)
    m, n, k = 64, 48, 576  # K of a 3x3 convolution over 64 channels
    a = rng.standard_normal((m, k)).astype(np.float32)
    b = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    ah, al = split(a)
    bh, bl = split(b)
    exact = (ah @ bh.T + ah @ bl.T + al @ bh.T).astype(np.float32)   # fp32 accumulation of the three tensor-core products
    fast = (ah @ bh.T).astype(np.float32)
    fp32 = a @ b.T
    e_exact, e_fast, e_fp32 = (np.abs(v - ref).max() for v in (exact, fast, fp32))
    assert e_exact < 4 * max(e_fp32, 1e-7)   # EXACT is as good as an fp32 GEMM (the dropped lo*lo term is ~2^-22 relative)
    assert e_fast > 50 * e_exact             # plain fp16 operands lose ~3 decimal digits against it
    assert e_fast < 5e-3
