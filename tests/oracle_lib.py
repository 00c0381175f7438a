"""ctypes loader for the CPU oracle (oracle/liboracle.so) and the compiled reference
(oracle/_ref/libref_rrtconnect.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_dp = C.POINTER(C.c_double)
c_u32p = C.POINTER(C.c_uint32)
c_i32p = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)


class NnfResult(C.Structure):
    _fields_ = [("n_vertices", C.c_uint32), ("n_edges", C.c_uint32), ("n_ik", C.c_uint32),
                ("lazy_iters", C.c_uint32), ("solved", C.c_int32), ("final_error", C.c_double),
                ("goal_cost", C.c_double), ("path_len", C.c_uint32), ("n_evaluated", C.c_uint32),
                ("n_collisions", C.c_uint32), ("overflow", C.c_uint32)]


class RefNnfResult(C.Structure):
    _fields_ = [("R", C.c_uint32), ("n_vertices", C.c_uint32), ("n_edges", C.c_uint32), ("n_ik", C.c_uint32),
                ("n_tpg_vertices", C.c_uint64), ("n_tpg_edges", C.c_uint64), ("solved", C.c_int32),
                ("final_error", C.c_double), ("goal_cost", C.c_double), ("path_len", C.c_uint32),
                ("n_evaluated", C.c_uint32), ("n_collisions", C.c_uint32), ("overflow", C.c_uint32),
                ("setup_seconds", C.c_double), ("solve_seconds", C.c_double)]


class RrtcResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("iterations", C.c_uint32), ("n_start", C.c_uint32),
                ("n_goal", C.c_uint32), ("path_len", C.c_uint32), ("overflow", C.c_uint32)]


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


_lib = None
_ref = None

_RRTC_ARGS = [C.c_void_p, C.c_int, c_dp, c_dp, c_dp, c_dp, C.c_double, C.c_int, C.c_int, C.c_double,
              C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(RrtcResult), c_dp,
              c_i32p, c_dp, c_i32p, c_dp]


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_knn.argtypes = [c_dp, C.c_uint64, C.c_int, c_dp, C.c_uint64, C.c_int, c_u32p, c_u32p, c_dp,
                              C.c_int]
        L.orc_kdtree_build.argtypes = [c_dp, C.c_uint64, C.c_int]
        L.orc_kdtree_build.restype = C.c_void_p
        L.orc_kdtree_destroy.argtypes = [C.c_void_p]
        L.orc_kdtree_knn.argtypes = [C.c_void_p, c_dp, C.c_uint64, C.c_int, c_u32p, c_dp, C.c_int]
        L.orc_distance.argtypes = [c_dp, c_dp, C.c_int]
        L.orc_distance.restype = C.c_double
        L.orc_interpolate.argtypes = [c_dp, c_dp, C.c_double, C.c_int, c_dp]
        L.orc_world_create.argtypes = [C.c_int, c_dp, C.c_int, c_i32p, c_dp, C.c_int, c_dp, C.c_int, c_dp]
        L.orc_world_create.restype = C.c_void_p
        L.orc_world_destroy.argtypes = [C.c_void_p]
        L.orc_clearance.argtypes = [C.c_void_p, c_dp]
        L.orc_clearance.restype = C.c_double
        L.orc_fk.argtypes = [C.c_void_p, c_dp, c_dp, c_dp]
        L.orc_check_motion_batch.argtypes = [C.c_void_p, c_dp, c_dp, C.c_uint64, C.c_double, C.c_int,
                                             C.c_int, c_u8p, c_dp, c_u32p, C.c_int]
        L.orc_is_valid_batch.argtypes = [C.c_void_p, c_dp, C.c_uint64, c_u8p, c_dp, C.c_int]
        L.orc_sample_uniform.argtypes = [C.c_uint64] * 4 + [C.c_double] * 2
        L.orc_sample_uniform.restype = C.c_double
        L.orc_rrtc_solve.argtypes = _RRTC_ARGS
        L.orc_lin_int_space.argtypes = [C.c_double, C.c_double, C.c_uint64, c_i32p]
        L.orc_nnf_layer_counts.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, c_u32p]
        L.orc_nnf_run.argtypes = [C.c_void_p, C.c_int, c_dp, C.c_int, C.c_int, C.c_double, c_u32p, c_dp,
                                  C.c_uint32, C.c_int, C.POINTER(NnfResult), c_u32p, c_dp, c_dp, C.c_uint32,
                                  c_u32p, C.c_uint32, c_dp]
        L.orc_nnf_reachable.argtypes = [C.c_int, c_dp, C.c_uint32, c_dp, C.c_uint32, c_u32p, c_u8p, C.c_double,
                                        C.c_int]
        _lib = L
    return _lib


def ref():
    """The reference's own RRTConnect.cpp compiled over ompl_min, or None if never built."""
    global _ref
    if _ref is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_rrtconnect.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.ref_rrtc_solve.argtypes = _RRTC_ARGS
        L.ref_set_two_phase.argtypes = [C.c_uint32, C.c_int]
        _ref = L
    return _ref


_ref_nnf = {}


def ref_nnf(reverse_order=False):
    """The reference's own frechet/src/*.cpp compiled over ompl_min + oracle/compat (mini-BGL), or None if
    never built.  reverse_order: the build whose stand-in hash_setS edge lists iterate backwards."""
    if reverse_order not in _ref_nnf:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_frechet_rev.so" if reverse_order else "libref_frechet.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.ref_nnf_run.argtypes = [C.c_void_p, C.c_uint32, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_dp,
                                  C.c_uint32, C.c_double, C.c_int, C.POINTER(RefNnfResult), c_dp, C.c_uint32,
                                  c_u32p, c_u32p, c_dp, c_dp, C.c_uint32, c_u32p, c_u8p, C.c_uint32, c_dp,
                                  C.c_uint32]
        L.ref_nnf_set_metric.argtypes = [C.c_int]
        _ref_nnf[reverse_order] = L
    return _ref_nnf[reverse_order]


def ref_nnf_run(world, ref_full_pose12, W, M, k, D, ik_states, seed=1, solve_seconds=600.0, search_only=False,
                reverse_order=False, metric=0):
    """NNFrechet::NNFrechet(si, W, M, k, D, seed) of the compiled reference: setRefPath(ref_full) ...
    setup(); solve(solve_seconds).  ik_states are handed out by the IK callback in call order."""
    L = ref_nnf(reverse_order)
    assert L is not None, "oracle/_ref/libref_frechet.so was never built"
    ref = np.ascontiguousarray(ref_full_pose12, dtype=np.float64).reshape(-1, 12)
    ik = np.ascontiguousarray(ik_states, dtype=np.float64).reshape(-1, world.dof)
    n_ik = len(ik)
    R = (W - 1) * (D + 1) + 1
    cap_v = 2 + n_ik + n_ik * k * D
    cap_e = 2 * M + n_ik * k * (D + 1) + 16
    cap_p = 4 * R + 64
    res = RefNnfResult()
    ref_sub = np.zeros((R, 12))
    calls = np.zeros(R, dtype=np.uint32)
    knn_out = np.empty((n_ik, k), dtype=np.uint32)
    vs = np.zeros((cap_v, world.dof))
    vp = np.zeros((cap_v, 12))
    edges = np.zeros((cap_e, 2), dtype=np.uint32)
    flags = np.zeros(cap_e, dtype=np.uint8)
    path = np.zeros((cap_p, world.dof))
    L.ref_nnf_set_metric(int(metric))  # 0 translation distance, 1 translation + rotation term
    try:
        rc = L.ref_nnf_run(world.h, len(ref), _dp(ref), W, M, k, D, seed, _dp(ik), n_ik, float(solve_seconds),
                           int(search_only), C.byref(res), _dp(ref_sub), R, calls.ctypes.data_as(c_u32p),
                           knn_out.ctypes.data_as(c_u32p), _dp(vs), _dp(vp), cap_v, edges.ctypes.data_as(c_u32p),
                           flags.ctypes.data_as(c_u8p), cap_e, _dp(path), cap_p)
    finally:
        L.ref_nnf_set_metric(0)
    assert rc == 0 and res.overflow == 0, (rc, res.overflow)
    out = {f[0]: getattr(res, f[0]) for f in RefNnfResult._fields_}
    out.update(ref=ref_sub[:res.R], ik_call_counts=calls[:res.R], knn=knn_out[:res.n_ik], states=vs[:res.n_vertices],
               pose=vp[:res.n_vertices], edges=edges[:res.n_edges], edge_flags=flags[:res.n_edges],
               path_states=path[:res.path_len].copy())
    return out


def knn(pts, queries, k, lo=None, nthreads=1):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    n, dim = pts.shape
    nq = queries.shape[0]
    idx = np.empty((nq, k), dtype=np.uint32)
    dist = np.empty((nq, k), dtype=np.float64)
    lo_p = None
    if lo is not None:
        lo = np.ascontiguousarray(lo, dtype=np.uint32)
        lo_p = lo.ctypes.data_as(c_u32p)
    rc = lib().orc_knn(_dp(pts), n, dim, _dp(queries), nq, k, lo_p, idx.ctypes.data_as(c_u32p), _dp(dist),
                       nthreads)
    assert rc == 0
    return idx, dist


class KdTree:
    """k-d tree over the same points with orc_knn's answers (the GNAT-class CPU baseline)."""

    def __init__(self, pts):
        self.pts = np.ascontiguousarray(pts, dtype=np.float64)
        self.dim = self.pts.shape[1]
        self.h = lib().orc_kdtree_build(_dp(self.pts), len(self.pts), self.dim)

    def knn(self, queries, k, nthreads=1):
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, self.dim)
        idx = np.empty((len(q), k), dtype=np.uint32)
        dist = np.empty((len(q), k), dtype=np.float64)
        rc = lib().orc_kdtree_knn(self.h, _dp(q), len(q), k, idx.ctypes.data_as(c_u32p), _dp(dist), nthreads)
        assert rc == 0
        return idx, dist

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_kdtree_destroy(self.h)
            self.h = None


class World:
    def __init__(self, dh, sphere_link, sphere_xyzr, obs_spheres, obs_boxes):
        self.dh = np.ascontiguousarray(dh, dtype=np.float64)
        self.link = np.ascontiguousarray(sphere_link, dtype=np.int32)
        self.xyzr = np.ascontiguousarray(sphere_xyzr, dtype=np.float64)
        self.os = np.ascontiguousarray(obs_spheres, dtype=np.float64).reshape(-1, 4)
        self.ob = np.ascontiguousarray(obs_boxes, dtype=np.float64).reshape(-1, 6)
        self.dof = self.dh.shape[0]
        self.h = lib().orc_world_create(self.dof, _dp(self.dh), len(self.link),
                                        self.link.ctypes.data_as(c_i32p), _dp(self.xyzr), len(self.os),
                                        _dp(self.os), len(self.ob), _dp(self.ob))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_world_destroy(self.h)
            self.h = None

    def clearance(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        return lib().orc_clearance(self.h, _dp(q))

    def fk(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        pose = np.empty(12)
        cen = np.empty((len(self.link), 3))
        lib().orc_fk(self.h, _dp(q), _dp(pose), _dp(cen))
        return pose.reshape(3, 4), cen

    def is_valid_batch(self, q, nthreads=1):
        q = np.ascontiguousarray(q, dtype=np.float64)
        n = q.shape[0]
        valid = np.empty(n, dtype=np.uint8)
        clear = np.empty(n, dtype=np.float64)
        lib().orc_is_valid_batch(self.h, _dp(q), n, valid.ctypes.data_as(c_u8p), _dp(clear), nthreads)
        return valid, clear

    def check_motion_batch(self, a, b, resolution, mode=0, check_first=False, want_clearance=True,
                           nthreads=1):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        n = a.shape[0]
        valid = np.empty(n, dtype=np.uint8)
        clear = np.empty(n, dtype=np.float64) if want_clearance else None
        ntest = np.empty(n, dtype=np.uint32) if want_clearance else None
        rc = lib().orc_check_motion_batch(self.h, _dp(a), _dp(b), n, resolution, mode, int(check_first),
                                          valid.ctypes.data_as(c_u8p), _dp(clear),
                                          ntest.ctypes.data_as(c_u32p) if ntest is not None else None,
                                          nthreads)
        assert rc == 0
        return valid, clear, ntest

    def rrtc_solve(self, start, goal, lo, hi, rng=0.0, ext=(0, 1), resolution_fraction=0.01, seed=0,
                   query=0, max_iters=1000, max_nodes=20000, max_path=4096, impl="port"):
        """impl: "port" (oracle/rrtconnect_port.cpp) or "ref" (the reference's RRTConnect.cpp)."""
        dim = self.dof
        start = np.ascontiguousarray(start, dtype=np.float64)
        goal = np.ascontiguousarray(goal, dtype=np.float64)
        lo = np.ascontiguousarray(np.broadcast_to(lo, (dim,)), dtype=np.float64)
        hi = np.ascontiguousarray(np.broadcast_to(hi, (dim,)), dtype=np.float64)
        res = RrtcResult()
        ss = np.zeros((max_nodes, dim))
        sp = np.zeros(max_nodes, dtype=np.int32)
        gs = np.zeros((max_nodes, dim))
        gp = np.zeros(max_nodes, dtype=np.int32)
        path = np.zeros((max_path, dim))
        fn = lib().orc_rrtc_solve if impl == "port" else ref().ref_rrtc_solve
        rc = fn(self.h, dim, _dp(lo), _dp(hi), _dp(start), _dp(goal), float(rng), int(ext[0]), int(ext[1]),
                float(resolution_fraction), int(seed), int(query), int(max_iters), int(max_nodes),
                int(max_path), C.byref(res), _dp(ss), sp.ctypes.data_as(c_i32p), _dp(gs),
                gp.ctypes.data_as(c_i32p), _dp(path))
        assert rc == 0 and res.overflow == 0, (rc, res.overflow)
        return dict(status=res.status, iterations=res.iterations, start_states=ss[:res.n_start].copy(),
                    start_parent=sp[:res.n_start].copy(), goal_states=gs[:res.n_goal].copy(),
                    goal_parent=gp[:res.n_goal].copy(), path=path[:res.path_len].copy())


def lin_int_space(mn, mx, n):
    out = np.empty(n, dtype=np.int32)
    lib().orc_lin_int_space(float(mn), float(mx), n, out.ctypes.data_as(c_i32p))
    return out


def nnf_layer_counts(seed, R, W, M):
    out = np.empty(R, dtype=np.uint32)
    lib().orc_nnf_layer_counts(seed, R, W, M, out.ctypes.data_as(c_u32p))
    return out


def nnf_reachable(ref_pose12, pose12, edges, tau, strict=False, blocked=None):
    """Independent bottleneck check: goal reachable through TPG cells with task distance <= tau (< if strict)?"""
    ref = np.ascontiguousarray(ref_pose12, dtype=np.float64).reshape(-1, 12)
    pose = np.ascontiguousarray(pose12, dtype=np.float64).reshape(-1, 12)
    ed = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, 2)
    bl = None
    if blocked is not None:
        blocked = np.ascontiguousarray(blocked, dtype=np.uint8)
        bl = blocked.ctypes.data_as(c_u8p)
    rc = lib().orc_nnf_reachable(len(ref), _dp(ref), len(pose), _dp(pose), len(ed), ed.ctypes.data_as(c_u32p), bl,
                                 float(tau), int(strict))
    assert rc in (0, 1), rc
    return bool(rc)


def nnf_run(world, ref_pose12, layer_counts, ik_states, k=5, D=3, check_res=0.05, max_lazy_iters=1000,
            nthreads=1, cap_vertices=None, cap_path=65536):
    ref = np.ascontiguousarray(ref_pose12, dtype=np.float64).reshape(-1, 12)
    counts = np.ascontiguousarray(layer_counts, dtype=np.uint32)
    ik = np.ascontiguousarray(ik_states, dtype=np.float64).reshape(-1, world.dof)
    n_ik = len(ik)
    if cap_vertices is None:
        cap_vertices = 2 + n_ik + n_ik * k * D
    res = NnfResult()
    knn_out = np.empty((n_ik, k), dtype=np.uint32)
    vs = np.zeros((cap_vertices, world.dof))
    vp = np.zeros((cap_vertices, 12))
    path = np.zeros(cap_path, dtype=np.uint32)
    trace = np.zeros(max_lazy_iters)
    rc = lib().orc_nnf_run(world.h, len(ref), _dp(ref), k, D, check_res, counts.ctypes.data_as(c_u32p), _dp(ik),
                           max_lazy_iters, nthreads, C.byref(res), knn_out.ctypes.data_as(c_u32p), _dp(vs), _dp(vp),
                           cap_vertices, path.ctypes.data_as(c_u32p), cap_path, _dp(trace))
    assert rc == 0 and res.overflow == 0, (rc, res.overflow)
    out = {f[0]: getattr(res, f[0]) for f in NnfResult._fields_}
    out.update(knn=knn_out, states=vs[:res.n_vertices], pose=vp[:res.n_vertices], path=path[:res.path_len].copy(),
               goal_cost_trace=trace[:res.lazy_iters].copy())
    return out
