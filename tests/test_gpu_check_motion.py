"""K2 parity: CUDA batched checkMotion / isValid / FK through the C ABI vs the CPU oracle.
Validity booleans must be identical except where the oracle's minimum |clearance| over the
tested set is below the 1e-6 contact band stated by BASELINE.json's north_star."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
BAND = 1e-6


def _compare(gv, ov, clear):
    diff = gv != ov
    assert not np.any(diff & (np.abs(clear) >= BAND)), np.argwhere(diff)[:10]
    return int(diff.sum())


def test_fk_matches(P, O, synth, gpu_world, oracle_world):
    q = synth.cloud(11, 500)
    g = gpu_world.fk_batch(q)
    for i in range(0, 500, 7):
        pose, _ = oracle_world.fk(q[i])
        assert np.allclose(g[i], pose, atol=1e-13)


def test_is_valid(P, O, synth, gpu_world, oracle_world):
    q = synth.cloud(12, 20000)
    gv = gpu_world.is_valid_batch(q)
    ov, oc = oracle_world.is_valid_batch(q, nthreads=8)
    assert _compare(gv, ov, oc) == 0
    assert 0.3 < gv.mean() < 0.9


@pytest.mark.parametrize("check_first", [False, True])
def test_check_motion_discrete(P, O, synth, gpu_world, oracle_world, check_first):
    a, b = synth.edges(2, 6000)
    gv = gpu_world.check_motion_batch(a, b, 0.01, check_first=check_first)
    ov, oc, nt = oracle_world.check_motion_batch(a, b, 0.01, check_first=check_first, nthreads=8)
    _compare(gv, ov, oc)
    assert 0.2 < gv.mean() < 0.8


def test_check_motion_edge_cases(P, O, synth, gpu_world, oracle_world):
    a, b = synth.edges(21, 512, max_len=0.02)       # nd in {0,1,2}: only endpoint(s) tested
    a[:64] = b[:64]                                  # zero-length edges
    a2, b2 = synth.edges(22, 256, max_len=6.0)      # long edges, many rounds per warp
    A, B = np.vstack([a, a2]), np.vstack([b, b2])
    for res in (0.01, 0.05, 0.5):
        gv = gpu_world.check_motion_batch(A, B, res)
        ov, oc, nt = oracle_world.check_motion_batch(A, B, res, nthreads=8)
        _compare(gv, ov, oc)
    assert gpu_world.check_motion_batch(A[:0], B[:0], 0.01).shape == (0,)


def test_check_motion_nnf_alpha(P, O, synth, gpu_world, oracle_world):
    # evaluateEdge's accumulated alpha: numQ values that skip alpha=1 (9, 11, 18, ...) are covered
    # by lengths spread over (0, 1.2]
    a, b = synth.edges(23, 4000, max_len=1.2)
    a[:16] = b[:16]
    gv = gpu_world.check_motion_batch(a, b, 0.05, mode=P.MODE_NNF_ALPHA)
    ov, oc, nt = oracle_world.check_motion_batch(a, b, 0.05, mode=1, nthreads=8)
    _compare(gv, ov, oc)


def test_free_world_all_valid(P, synth, world_arrays):
    dh, link, xyzr, _, _ = world_arrays
    w = P.World(dh, link, xyzr, np.zeros((0, 4)), np.zeros((0, 6)))
    a, b = synth.edges(5, 300)
    assert np.all(w.check_motion_batch(a, b, 0.01) == 1)


@pytest.mark.parametrize("mode", [0, 1])
def test_thin_obstacles_conservative_advancement(P, O, synth, world_arrays, mode):
    # Obstacles far thinner than the spacing of phase A's samples: an edge is typically invalid at one or
    # two isolated states only, which conservative advancement (states struck off because a neighbouring
    # sample has enough clearance) must never skip.  Few obstacles -> large clearances -> long reaches.
    dh, link, xyzr, _, _ = world_arrays
    rng = np.random.default_rng(31)
    n_s, n_b = 40, 24
    d = rng.normal(size=(n_s, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    sph = np.hstack([d * rng.uniform(0.25, 1.1, (n_s, 1)), rng.uniform(0.003, 0.012, (n_s, 1))])
    sph[:, 2] = np.abs(sph[:, 2]) + 0.15
    d = rng.normal(size=(n_b, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    half = np.full((n_b, 3), 0.15); half[np.arange(n_b), rng.integers(0, 3, n_b)] = 0.002   # thin plates
    box = np.hstack([d * rng.uniform(0.4, 1.1, (n_b, 1)), half])
    box[:, 2] = np.abs(box[:, 2]) + 0.25
    gw = P.World(dh, link, xyzr, sph, box)
    ow = O.World(dh, link, xyzr, sph, box)
    for seed, n, max_len in ((41, 9000, 1.0), (42, 5000, 3.0), (43, 1500, 1.0)):   # flattened path twice, warp path once
        a, b = synth.edges(seed, n, max_len=max_len)
        res = 0.01 if mode == 0 else 0.02
        gv = gw.check_motion_batch(a, b, res, mode=mode)
        ov, oc, nt = ow.check_motion_batch(a, b, res, mode=mode, nthreads=8)
        _compare(gv, ov, oc)
        assert 0.05 < gv.mean() < 0.98
    # states: the field + lists against the oracle
    q = synth.cloud(44, 30000)
    ov, oc = ow.is_valid_batch(q, nthreads=8)
    assert _compare(gw.is_valid_batch(q), ov, oc) == 0


def test_field_and_grid_switches_do_not_change_verdicts(P, synth, world_arrays, monkeypatch):
    # the clearance field (+ conservative advancement) and the obstacle grid are accelerations only
    a, b = synth.edges(61, 20000, max_len=1.5)
    q = synth.cloud(62, 20000)
    ref_v = ref_s = None
    for env in ({}, {"PRGPU_NO_DFIELD": "1"}, {"PRGPU_NO_GRID": "1"}):
        for k in ("PRGPU_NO_DFIELD", "PRGPU_NO_GRID"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        w = P.World(*world_arrays)
        v = w.check_motion_batch(a, b, 0.01)
        s = w.is_valid_batch(q)
        if ref_v is None:
            ref_v, ref_s = v, s
        else:
            assert np.array_equal(v, ref_v) and np.array_equal(s, ref_s), env


def test_item_table_overflow_changes_nothing(P, synth, gpu_world, monkeypatch):
    """Large batches deal the obstacle-list walks out as (state, sphere) items; when the item table is full a lane
    walks its own lists on the spot.  With the table shrunk to almost nothing the verdicts must be the same."""
    a, b = synth.edges(2, 60000)
    ref = gpu_world.check_motion_batch(a, b, 0.01)
    ref_nnf = gpu_world.check_motion_batch(a, b, 0.05, mode=P.MODE_NNF_ALPHA)
    for cap in ("0", "7", "1000"):
        monkeypatch.setenv("PRGPU_CM_ITEM_CAP", cap)
        assert np.array_equal(gpu_world.check_motion_batch(a, b, 0.01), ref), cap
        assert np.array_equal(gpu_world.check_motion_batch(a, b, 0.05, mode=P.MODE_NNF_ALPHA), ref_nnf), cap
    monkeypatch.delenv("PRGPU_CM_ITEM_CAP")
    assert np.array_equal(gpu_world.check_motion_batch(a, b, 0.01), ref)


@pytest.mark.parametrize("mode_name", ["MODE_OMPL_DISCRETE", "MODE_NNF_ALPHA"])
def test_single_edge_calls_through_the_host_window(P, synth, gpu_world, monkeypatch, mode_name):
    """A planner's checkMotion() is one edge per call: the edge and the verdict travel through a mapped host window.
    One edge at a time, 7 at a time and the copy path (PRGPU_NO_HOST_WINDOW) all give the batch's verdicts."""
    mode = getattr(P, mode_name)
    a, b = synth.edges(91, 120, max_len=1.2)
    ref = gpu_world.check_motion_batch(a, b, 0.02, mode=mode)       # 120 edges: the copy path
    one = np.concatenate([gpu_world.check_motion_batch(a[i:i + 1], b[i:i + 1], 0.02, mode=mode) for i in range(120)])
    seven = np.concatenate([gpu_world.check_motion_batch(a[i:i + 7], b[i:i + 7], 0.02, mode=mode) for i in range(0, 120, 7)])
    assert np.array_equal(one, ref) and np.array_equal(seven, ref)
    monkeypatch.setenv("PRGPU_NO_HOST_WINDOW", "1")
    one = np.concatenate([gpu_world.check_motion_batch(a[i:i + 1], b[i:i + 1], 0.02, mode=mode) for i in range(120)])
    assert np.array_equal(one, ref)
