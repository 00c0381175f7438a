"""The error bounds the kNN filter's proof relies on (csrc/knn_filter.cu), checked on the CPU with numpy.

eps_tc bounds |D - (|x|^2 - 2 q.x - T)| for the tensor-core scan: both factors rounded fp64 -> fp32 -> TF32
(round to nearest, ties away), products exact, fp32 accumulation (modelled here as one fp32 rounding per
add in the worst order: the hardware is at least that precise, and the bound carries a 32x allowance for
it).  eps_32 bounds the fp32 FMA chain of the sample / re-cut passes.  Not a kernel test: it pins the
formulas, so that a change of either side shows up without a GPU."""
import numpy as np


def to_tf32(a):
    """fp32 -> TF32 (10-bit mantissa), round to nearest, ties away from zero (cvt.rna.tf32.f32)."""
    b = np.asarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x1000) & 0xFFFFE000
    return b.astype(np.uint32).view(np.float32)


def eps_tc(qn, xn, T):
    return 1.01 * 2.0 ** -9 * qn * xn + 2.0 ** -18 * ((xn + qn) ** 2 + abs(T)) + 1e-30


def eps_32(qn, xn):
    return 16.0 * 2.0 ** -24 * (xn + qn) ** 2 * (1.0 + 1e-6) + 1e-30


def test_tf32_rounding_helper():
    x = np.float32(1.0) + np.float32(2.0 ** -11)          # exactly half way between two TF32 numbers
    assert to_tf32(x) == np.float32(1.0 + 2.0 ** -10)     # ties away
    assert to_tf32(np.float32(-3.0)) == np.float32(-3.0)
    r = np.random.default_rng(0).normal(size=10000).astype(np.float32)
    assert np.all(np.abs(to_tf32(r).astype(np.float64) - r) <= np.abs(r) * 2.0 ** -11)


def test_scan_error_within_eps_tc():
    rng = np.random.default_rng(1)
    for scale, offset in ((1.0, 0.0), (1e-3, 0.0), (1e3, 0.0), (0.01, 250.0), (1.0, -40.0)):
        x = rng.uniform(-np.pi, np.pi, (20000, 7)) * scale + offset
        q = rng.uniform(-np.pi, np.pi, (64, 7)) * scale + offset
        xn = np.sqrt((x * x).sum(1).max())
        xt = to_tf32(x.astype(np.float32)).astype(np.float64)             # the TF32 mirror
        n2 = (x * x).sum(1).astype(np.float32)                           # |x|^2 kept in fp32
        for qi in q:
            qn = np.sqrt((qi * qi).sum())
            a = to_tf32((-2.0 * qi).astype(np.float32)).astype(np.float64)
            key_real = (x * x).sum(1) - 2.0 * x @ qi
            T = np.float32(np.partition(key_real, 512)[512])
            T = np.float32(np.nextafter(to_tf32(T), np.float32(np.inf))) if to_tf32(T) < T else to_tf32(T)
            # products exact (float64 holds 11x11-bit products exactly), worst-order fp32 accumulation
            acc = n2.astype(np.float32)
            for d in range(7):
                acc = (acc.astype(np.float64) + a[d] * xt[:, d]).astype(np.float32)
            D = (acc.astype(np.float64) - np.float64(T)).astype(np.float32)
            err = np.abs(D.astype(np.float64) - (key_real - np.float64(T)))
            assert err.max() <= eps_tc(qn, xn, float(T)), (scale, offset, err.max(), eps_tc(qn, xn, float(T)))


def test_fp32_key_error_within_eps_32():
    rng = np.random.default_rng(2)
    for scale, offset in ((1.0, 0.0), (1e3, 5e3), (1e-2, 1.0)):
        x = rng.uniform(-np.pi, np.pi, (20000, 7)) * scale + offset
        q = rng.uniform(-np.pi, np.pi, (32, 7)) * scale + offset
        xn = np.sqrt((x * x).sum(1).max())
        xf = x.astype(np.float32)
        n2 = (x * x).sum(1).astype(np.float32)
        for qi in q:
            qn = np.sqrt((qi * qi).sum())
            q2 = (-2.0 * qi).astype(np.float32)
            acc = n2.copy()
            for d in range(7):   # fmaf(q2[d], x[d], acc): one rounding per step
                acc = (q2[d].astype(np.float64) * xf[:, d].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
            key_real = (x * x).sum(1) - 2.0 * x @ qi
            err = np.abs(acc.astype(np.float64) - key_real)
            assert err.max() <= eps_32(qn, xn), (scale, offset, err.max(), eps_32(qn, xn))
