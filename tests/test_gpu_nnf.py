"""K3 parity: NNF on the device (suffix-masked kNN, sub-sampled edges + FK, wavefront bottleneck DP,
LazySP) vs the CPU oracle port of frechet/src/{NNFrechet,LPAStar,util}.cpp.  kNN choices and
interpolated configurations are bit-exact; poses agree to 1e-12 (sincos ulps); Frechet / goal costs
must agree within 1e-5 relative (north star) -- in practice to ~1e-15."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nnf_worlds(P, O, synth):
    dh, link, xyzr = synth.default_arm()
    sph, box = synth.obstacles(3, n_spheres=32, n_boxes=32)
    return dh, P.World(dh, link, xyzr, sph, box), O.World(dh, link, xyzr, sph, box)


def _problem(O, synth, dh, seed, W, M, D):
    R = (W - 1) * (D + 1) + 1                       # subsampleRefPath, NNFrechet.cpp:237
    q, poses = synth.nnf_reference(seed, 2 * R + 5, dh)
    ids = O.lin_int_space(0, len(q) - 1, R)
    counts = O.nnf_layer_counts(1, R, W, M)
    assert counts.sum() == W * M and counts[0] == M and counts[-1] == M
    ik = synth.nnf_ik_states(2, q[ids], counts)
    return poses[ids], counts, ik


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("seed,W,M,k,D", [(4, 5, 5, 5, 3), (3, 12, 20, 5, 1), (4, 20, 30, 4, 2), (1, 12, 20, 5, 1),
                                          (2, 8, 6, 3, 1), (4, 30, 40, 5, 1)])
def test_nnf_matches_oracle(P, O, synth, nnf_worlds, seed, W, M, k, D, mode):
    dh, gw, ow = nnf_worlds
    ref, counts, ik = _problem(O, synth, dh, seed, W, M, D)
    o = O.nnf_run(ow, ref, counts, ik, k=k, D=D, max_lazy_iters=3000, nthreads=8)
    g = P.Nnf(gw, ref, counts, ik, num_nn=k, discretization=D)
    g.set_mode(mode)  # 0: banded column-wise DP, 1: full-table wavefront
    assert (g.n_vertices, g.n_edges, g.n_ik) == (o["n_vertices"], o["n_edges"], o["n_ik"])
    assert np.array_equal(g.knn(), o["knn"])                      # findKnnNodes: bit-exact choices
    st, po = g.vertices()
    assert np.array_equal(st, o["states"])                        # addSubsampledEdge lerp: bit-exact
    assert np.allclose(po, o["pose"], rtol=0, atol=1e-12)
    r = g.solve(max_lazy_iters=3000)
    assert r["solved"] == o["solved"]
    assert r["lazy_iters"] == o["lazy_iters"]
    assert (r["n_evaluated"], r["n_collisions"]) == (o["n_evaluated"], o["n_collisions"])
    assert np.array_equal(r["path"], o["path"])
    tr_g, tr_o = r["goal_cost_trace"], o["goal_cost_trace"]
    fin = tr_o < 1e300
    assert np.array_equal(fin, tr_g < 1e300)
    assert np.allclose(tr_g[fin], tr_o[fin], rtol=1e-5, atol=0)   # north-star tolerance
    assert np.allclose(tr_g[fin], tr_o[fin], rtol=1e-12, atol=0)  # what is actually achieved
    if o["lazy_iters"]:
        assert abs(r["final_error"] - o["final_error"]) <= 1e-5 * max(o["final_error"], 1e-30)
    assert g.last_dp_ms() > 0.0


def test_nnf_free_world_first_path_is_solution(P, O, synth):
    dh, link, xyzr = synth.default_arm()
    gw = P.World(dh, link, xyzr, np.zeros((0, 4)), np.zeros((0, 6)))
    ow = O.World(dh, link, xyzr, np.zeros((0, 4)), np.zeros((0, 6)))
    ref, counts, ik = _problem(O, synth, dh, 5, 10, 16, 3)
    g = P.Nnf(gw, ref, counts, ik, num_nn=5, discretization=3)
    r = g.solve()
    o = O.nnf_run(ow, ref, counts, ik, k=5, D=3)
    assert r["solved"] == 1 and r["lazy_iters"] == 1 and r["n_collisions"] == 0
    assert np.array_equal(r["path"], o["path"])
    assert abs(r["final_error"] - o["final_error"]) <= 1e-12
    # the bottleneck cost can only be improved by denser sampling, never beaten by any other path:
    # re-solving is idempotent
    r2 = g.solve()
    assert r2["lazy_iters"] == 1 and np.array_equal(r2["path"], r["path"])


def test_nnf_banded_equals_full_at_scale(P, O, synth, nnf_worlds):
    """Mid-size problem (too slow for the CPU oracle's LazySP): both device searches must agree on
    every LazySP iteration's goal cost, the evaluated/colliding edge counts and the final path."""
    dh, gw, _ = nnf_worlds
    ref, counts, ik = _problem(O, synth, dh, 4, 40, 128, 1)
    out = []
    for mode in (0, 1, 2):
        g = P.Nnf(gw, ref, counts, ik, num_nn=5, discretization=1)
        g.set_mode(mode)
        out.append(g.solve(max_lazy_iters=60))
        if mode != 1:
            info = g.band_info()
            assert 0 < info["cells"] < len(ref) * g.n_vertices   # the band is a strict subset
    a = out[0]
    for b in out[1:]:
        assert np.array_equal(a["goal_cost_trace"], b["goal_cost_trace"])
        assert (a["lazy_iters"], a["n_evaluated"], a["n_collisions"], a["solved"]) == \
               (b["lazy_iters"], b["n_evaluated"], b["n_collisions"], b["solved"])
        assert np.array_equal(a["path"], b["path"]) and a["final_error"] == b["final_error"]


def test_nnf_bad_params(P, synth, nnf_worlds):
    dh, gw, _ = nnf_worlds
    ref = np.zeros((3, 12))
    with pytest.raises(P.PrgpuError):
        P.Nnf(gw, ref, [1, 1, 1], np.zeros((3, 7)), discretization=0)   # "discretization must be positive!"
    with pytest.raises(P.PrgpuError):
        P.Nnf(gw, ref, [1, 1, 1], np.zeros((3, 7)), num_nn=0)            # "numNN must be positive!"
    with pytest.raises(P.PrgpuError):
        P.Nnf(gw, ref[:1], [1], np.zeros((1, 7)))                        # "referencePath is empty!"
