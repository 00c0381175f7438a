"""The error bounds the exactness proofs rest on, measured ON THE DEVICE.

K1: knn_select_kernel proves a query from  |D - (|x|^2 - 2 q.x - T)| <= eps_tc  for every node the tensor-core
scan looked at, and re-cuts with  |key32 - (|x|^2 - 2 q.x)| <= eps_32.  prgpu_diag_knn_scan runs the PRODUCTION
scan kernel and returns its D, the fp32 keys and the two bounds as the kernel computes them; here they are held
against float64 / longdouble arithmetic over the adversarial families of test_filter_path_adversarial_ties.
K2: DeviceWorld::delta must bound the distance between the production fp32 FK chain's sphere centres and the
fp64 chain's (prgpu_diag_fk_centre_error), over 1e6 states."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _families(synth):
    rng = np.random.default_rng(3)
    yield "uniform", synth.cloud(0, 20000), synth.cloud(1, 192)
    base = synth.cloud(6, 2500)
    yield "duplicates", np.repeat(base, 8, axis=0)[rng.permutation(20000)], np.vstack([base[:100], synth.cloud(7, 92)])
    centres = synth.cloud(8, 1000)
    pts = (centres[:, None, :] + rng.uniform(-1, 1, (1000, 20, 7)) * 10.0 ** rng.uniform(-9, -6, (1000, 20, 1))
           ).reshape(-1, 7)
    yield "near-duplicates", pts, np.vstack([centres[:100], synth.cloud(9, 92)])
    yield "far-from-origin", synth.cloud(10, 20000) * 0.01 + 250.0, synth.cloud(11, 192) * 0.01 + 250.0
    centres = synth.cloud(12, 64)
    pts = (centres[:, None, :] + rng.normal(0.0, 0.02, (64, 300, 7))).reshape(-1, 7)
    yield "tight-blobs", pts, np.vstack([centres + rng.normal(0.0, 0.02, centres.shape), synth.cloud(13, 128)])
    for scale in (1e6, 1e-6):
        yield "scale%g" % scale, synth.cloud(14, 20000) * scale, synth.cloud(15, 192) * scale
    # not fp32-representable coordinates (Gaussian IK states of the NNF workload)
    yield "gaussian", rng.normal(0.0, 1.3, (20000, 7)), rng.normal(0.0, 1.3, (192, 7))


def _exact_keys(pts, q):
    """|x|^2 - 2 q.x in extended precision, [nq, n]"""
    x = pts.astype(np.longdouble)
    n2 = (x * x).sum(1)
    return (n2[None, :] - 2.0 * (q.astype(np.longdouble) @ x.T)).astype(np.float64)


def test_scan_values_within_eps_tc_on_device(P, synth):
    worst_tc, worst_32 = 0.0, 0.0
    for name, pts, q in _families(synth):
        n, nq = len(pts), len(q)
        t = P.KnnIndex(7)
        t.add(pts)
        key = _exact_keys(pts, q)
        # (1) a threshold above every key: the scan reports EVERY node, so every D it can produce is checked
        T_all = np.float32(key.max() * 1.001 + abs(key).max() * 1e-3 + 1e-30)
        r = t.diag_scan(q, T_all, cap=n)
        assert (r["counts"] == n).all(), (name, r["counts"].min(), n)
        for i in range(nq):
            idx = r["idx"][i]
            assert np.array_equal(np.sort(idx), np.arange(n))
            exact = key[i, idx]
            err_tc = np.abs(r["D"][i].astype(np.float64) - (exact - np.float64(r["thresh"][i])))
            err_32 = np.abs(r["key32"][i].astype(np.float64) - exact)
            assert err_tc.max() <= r["eps_tc"][i], (name, i, err_tc.max(), r["eps_tc"][i])
            assert err_32.max() <= r["eps_32"][i], (name, i, err_32.max(), r["eps_32"][i])
            worst_tc = max(worst_tc, err_tc.max() / r["eps_tc"][i])
            worst_32 = max(worst_32, err_32.max() / r["eps_32"][i])
        # (2) a realistic threshold (about the 300th smallest key): the proof's actual claim -- no node whose
        #     exact key lies below T - eps_tc is missing from the report, and nothing reported lies above T + eps_tc
        T_q = np.partition(key, 300, axis=1)[:, 300].astype(np.float32)
        r = t.diag_scan(q, T_q, cap=n)     # (fp32 cannot resolve T among keys of magnitude 1e5..1e13: counts up to n)
        for i in range(nq):
            c = int(r["counts"][i])
            assert c <= n
            rep = np.zeros(n, dtype=bool)
            rep[r["idx"][i, :c]] = True
            Tr, e = np.float64(r["thresh"][i]), r["eps_tc"][i]
            assert Tr >= np.float64(T_q[i])                       # rounded UP to a TF32 number
            assert rep[key[i] < Tr - e].all(), (name, i)
            assert (key[i, rep] <= Tr + e).all(), (name, i)
            err = np.abs(r["D"][i, :c].astype(np.float64) - (key[i, r["idx"][i, :c]] - Tr))
            assert err.max(initial=0.0) <= e, (name, i, err.max(), e)
        t.close()
    # the bounds are not vacuous either: the worst observed error is a visible fraction of them
    assert 1e-3 < worst_tc <= 1.0 and 1e-3 < worst_32 <= 1.0, (worst_tc, worst_32)


def test_fp32_fk_centres_within_delta_on_device(P, synth, gpu_world):
    rng = np.random.default_rng(5)
    worst = 0.0
    for lo, hi, n in ((-np.pi, np.pi, 700_000), (-2 * np.pi, 2 * np.pi, 200_000), (-1e-3, 1e-3, 50_000)):
        q = rng.uniform(lo, hi, (n, 7))
        err, delta = gpu_world.diag_fk_centre_error(q)
        assert delta > 0 and np.isfinite(err).all()
        assert err.max() <= delta / 4.0, (lo, hi, err.max(), delta)
        worst = max(worst, err.max())
    # joint values at the exact multiples of pi/2 (sincosf's reduction boundaries)
    q = rng.integers(-4, 5, (50_000, 7)) * (np.pi / 2)
    err, delta = gpu_world.diag_fk_centre_error(q)
    assert err.max() <= delta / 4.0
    worst = max(worst, err.max())
    assert worst > delta * 1e-4      # a bound within 4 orders of the observed error, not an arbitrary constant
