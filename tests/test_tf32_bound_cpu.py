"""
CPU check of the error model behind the certificate of the tensor-core exact top-k (csrc/exact_tf32_kernels.cu):
the 3xTF32 dot product  sum_k (lo*hi + hi*lo + hi*hi)  with hi = tf32(x), lo = tf32(x - hi) and one TRUNCATING fp32
accumulation per MMA (the pessimistic model of the hardware accumulator) must stay within
tf32_dot_error(dim) * |q| |x| = (0.5625 dim + 12) * 2^-22 * |q| |x|  of the exact dot product -- with a wide margin,
because the kernel relies on that bound to prove that no vector outside its candidates belongs to the top k.
"""
import numpy as np


def _tf32(x):
    """cvt.rna.tf32.f32: keep 10 mantissa bits, round to nearest, ties away from zero"""
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x1000) & 0xFFFFE000
    return b.astype(np.uint32).view(np.float32)


def _trunc32(v):
    """float64 -> float32, rounded toward zero"""
    f = v.astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(v)
    return np.where(over, np.nextafter(f, np.float32(0)), f).astype(np.float32)


def _dot_3xtf32(q, x):
    """q [n, dim], x [n, dim] -> [n]; K-steps of 8, three MMAs per step, truncating fp32 accumulator"""
    qh, xh = _tf32(q), _tf32(x)
    ql, xl = _tf32(q - qh), _tf32(x - xh)
    acc = np.zeros(q.shape[0], dtype=np.float32)
    for k0 in range(0, q.shape[1], 8):
        s = slice(k0, k0 + 8)
        for a, b in ((ql, xh), (qh, xl), (qh, xh)):
            prod = (a[:, s].astype(np.float64) * b[:, s].astype(np.float64)).sum(axis=1)  # exact: 11 x 11 bit products
            acc = _trunc32(acc.astype(np.float64) + prod)
    return acc


def _bound(dim):
    return (0.5625 * dim + 12.0) * 2.0 ** -22


def test_tf32_split_is_exact_and_small():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(100000) * np.exp(rng.uniform(-20, 20, 100000))).astype(np.float32)
    hi = _tf32(x)
    lo = (x - hi).astype(np.float32)
    assert np.array_equal(hi.astype(np.float64) + lo.astype(np.float64), x.astype(np.float64))  # x - hi is exact in fp32
    assert (np.abs(lo) <= np.abs(x) * 2.0 ** -11 * (1 + 1e-6)).all()
    assert (np.abs(lo - _tf32(lo)) <= np.abs(x) * 2.0 ** -22 * (1 + 1e-6)).all()


def test_3xtf32_dot_error_stays_far_inside_the_certificate_bound():
    rng = np.random.default_rng(1)
    worst = 0.0
    for dim in (8, 64, 96, 128, 200, 960):
        n = 4000
        cases = [
            (rng.standard_normal((n, dim)), rng.standard_normal((n, dim))),                       # generic
            (rng.random((n, dim)), rng.random((n, dim))),                                          # all positive: no cancellation
            (rng.standard_normal((n, dim)) * np.exp(rng.uniform(-6, 6, (n, dim))),               # wild dynamic range
             rng.standard_normal((n, dim)) * np.exp(rng.uniform(-6, 6, (n, dim)))),
        ]
        v = rng.standard_normal((n, dim))
        cases.append((v, v + 1e-3 * rng.standard_normal((n, dim))))                                 # near-identical vectors (large |q||x|)
        for q, x in cases:
            q = q.astype(np.float32)
            x = x.astype(np.float32)
            exact = (q.astype(np.float64) * x.astype(np.float64)).sum(axis=1)
            scale = np.linalg.norm(q.astype(np.float64), axis=1) * np.linalg.norm(x.astype(np.float64), axis=1)
            err = np.abs(_dot_3xtf32(q, x).astype(np.float64) - exact) / scale
            assert err.max() <= 0.5 * _bound(dim), (dim, err.max(), _bound(dim))
            worst = max(worst, err.max() / _bound(dim))
    assert worst > 0  # (the model is exercised)
