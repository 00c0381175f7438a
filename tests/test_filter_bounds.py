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


# ---- the fp32 leaf filter of the k-d index (csrc/kdtree.cu: kd_T32) ------------------------------------------
def kd_err(qn, mn):
    return (qn + mn) * (1.0 + 2.0 ** -40) * (2.0 ** -24 * (1.0 + 2.0 ** -23)) + 1e-40


def kd_T32(r, E):
    t = r * (1.0 + 2.0 ** -21) + E
    t = t * t * (1.0 + 2.0 ** -19) + 1e-37
    with np.errstate(over="ignore"):
        f = np.float32(t)
    if np.float64(f) < t:                                  # __double2float_ru
        f = np.nextafter(f, np.float32(np.inf))
    return f


def fp32_chain(x32, q32):
    """s32 of the leaf pass: a = x32 - q32 (fp32), s = fma(a, a, s) in dimension order."""
    with np.errstate(over="ignore", invalid="ignore"):
        a = (x32 - q32).astype(np.float32)
        s = (a[:, 0].astype(np.float64) * a[:, 0]).astype(np.float32)
        for d in range(1, x32.shape[1]):
            s = (a[:, d].astype(np.float64) * a[:, d] + s.astype(np.float64)).astype(np.float32)
    return s


def test_kd_leaf_filter_never_rejects_a_node_within_the_bound():
    """The filter rejects a node iff s32 > T32(r).  So every node whose fp64 distance (as the reference computes
    it) is <= r must have s32 <= T32(r) -- for r the distance of the node itself (the tightest r it can meet)."""
    rng = np.random.default_rng(7)
    cases = [(1.0, 0.0), (1e-3, 0.0), (1e6, 0.0), (1e-6, 0.0), (1e-6, 250.0), (1e-9, -40.0), (1.0, 1e5),
             (1e-30, 0.0), (1e-42, 0.0), (1e30, 0.0)]
    for scale, offset in cases:
        x = rng.uniform(-np.pi, np.pi, (30000, 7)) * scale + offset
        q = rng.uniform(-np.pi, np.pi, (32, 7)) * scale + offset
        q[:8] = x[:8]                                       # queries that are nodes
        q[8:16] = x[8:16] + 1e-7 * scale                    # and almost nodes
        mn = np.sqrt((np.abs(x).max(0) ** 2).sum())         # M: largest norm inside the bounding box
        with np.errstate(over="ignore"):
            x32 = x.astype(np.float32)
        for qi in q:
            E = kd_err(np.sqrt((qi * qi).sum()), mn)
            with np.errstate(over="ignore"):
                s32 = fp32_chain(x32, qi.astype(np.float32)[None, :])
            a = x - qi
            s = a[:, 0] * a[:, 0]
            for d in range(1, 7):
                s = s + a[:, d] * a[:, d]                   # the reference's order: sub, mul, add
            dist = np.sqrt(s)
            T = np.array([kd_T32(r, E) for r in dist[:2000]], dtype=np.float32)
            assert np.all(~(s32[:2000] > T)), (scale, offset)
            # and the bound is not vacuous: at the scale of the data it rejects what is 1e-3 farther
            if offset == 0.0 and 1e-6 <= scale <= 1e6:
                far = dist[:2000] * (1 - 1e-3)
                Tf = np.array([kd_T32(r, E) for r in far], dtype=np.float32)
                assert np.mean(s32[:2000] > Tf) > 0.95, (scale, offset)


def test_kd_leaf_filter_out_of_range_coordinates_pass():
    """Coordinates beyond fp32 become inf, differences inf or NaN: never rejected against an infinite threshold, and
    NaN is never rejected at all (the kernel tests !(s32 > T32))."""
    x32 = np.array([[np.inf, 0, 0, 0, 0, 0, 0], [np.inf, 0, 0, 0, 0, 0, 0]], dtype=np.float32)
    q32 = np.array([[np.inf, 0, 0, 0, 0, 0, 0]], dtype=np.float32)
    s = fp32_chain(x32, q32)
    assert np.all(np.isnan(s)) and not np.any(s > np.float32(1.0))
    assert kd_T32(1e30, 0.0) == np.float32(np.inf)          # a threshold beyond fp32 rejects nothing
    assert not (np.float32(np.inf) > kd_T32(1e30, 0.0))
