"""The operand split of the tensor-core layers (kernels.cuh split_f16, gemm_tc.cuh), emulated in numpy: x = x1 + x2s * 2^-11 with
x1 = fp16(x), x2s = fp16((x - x1) * 2^11).  Checks the error bounds DESIGN.md quotes and that a three-product fp32-accumulated dot
product (main = a1.b1, corr = a2s.b1 + a1.b2s) is fp32-grade."""
import numpy as np


def split(x):
    x = x.astype(np.float32)
    x1 = x.astype(np.float16)
    x2s = ((x - x1.astype(np.float32)) * np.float32(2048)).astype(np.float16)
    return x1, x2s


def test_reconstruction_error_bounds():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200000) * np.exp(rng.uniform(-6, 6, 200000))).astype(np.float32)      # |x| from ~1e-3 to ~1e3
    x1, x2s = split(x)
    rec = x1.astype(np.float64) + x2s.astype(np.float64) / 2048
    xa = np.abs(x.astype(np.float64))
    normal = xa >= 2.0 ** -14                           # fp16's normal range; below it the bound is absolute (next assertion)
    assert normal.mean() > 0.9
    assert (np.abs(rec - x.astype(np.float64))[normal] / xa[normal]).max() <= 2.0 ** -22      # 22 significand bits (fp32 has 24)
    assert np.abs(rec - x.astype(np.float64))[~normal].max() <= 2.0 ** -35
    tiny = (rng.standard_normal(10000) * 1e-7).astype(np.float32)                                   # below fp16's normal range
    t1, t2 = split(tiny)
    assert np.abs(t1.astype(np.float64) + t2.astype(np.float64) / 2048 - tiny).max() <= 2.0 ** -35  # absolute, not relative
    big = np.array([65504.0, -65504.0, 70000.0], np.float32)
    with np.errstate(over='ignore'):
        b1, _ = split(big)
    assert np.isfinite(b1[:2].astype(np.float32)).all() and not np.isfinite(b1[2].astype(np.float32))   # > 65504: range fallback


def test_three_product_dot_is_fp32_grade():
    rng = np.random.default_rng(1)
    K = 768
    a = rng.standard_normal((64, K)).astype(np.float32)
    b = (rng.standard_normal((32, K)) / np.sqrt(K)).astype(np.float32)
    a1, a2 = split(a); b1, b2 = split(b)
    f = lambda h: h.astype(np.float32)
    main = f(a1) @ f(b1).T
    corr = f(a2) @ f(b1).T + f(a1) @ f(b2).T
    got = main + corr * np.float32(1 / 2048)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    err = np.abs(got - ref).max() / np.abs(ref).max()
    single = np.abs(f(a1) @ f(b1).T - ref).max() / np.abs(ref).max()
    assert err < 2e-6 and single > 50 * err            # one fp16 pass alone is ~1e-3: the reason for the split
