"""ctypes loader of libprgpu_plugin.so (the C++ OMPL-plugin mirror and its C test driver)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
c_vp = C.c_void_p


class PlgResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("iterations", C.c_uint32), ("n_start", C.c_uint32),
                ("n_goal", C.c_uint32), ("path_len", C.c_uint32), ("overflow", C.c_uint32)]


class PlgNnfResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("lazy_iters", C.c_uint32), ("n_evaluated", C.c_uint32),
                ("n_collisions", C.c_uint32), ("path_len", C.c_uint32), ("R", C.c_uint32),
                ("final_error", C.c_double), ("goal_cost", C.c_double)]


_lib = None
_dropin = None
DROPIN_PATH = os.path.join(ROOT, "oracle", "_ref", "libdropin_a.so")


def dropin():
    """oracle/_ref/libdropin_a.so: the reference's own RRTConnect.cpp, compiled unmodified, over the product's
    GpuNearestNeighbors / GpuStateValidityChecker / GpuMotionValidator ("Drop-in A").  None if never built."""
    global _dropin
    if _dropin is None:
        if not os.path.exists(DROPIN_PATH):
            return None
        L = C.CDLL(DROPIN_PATH)
        L.plg_world_create.restype = c_vp
        L.plg_world_create.argtypes = [C.c_int, C.c_int, c_vp, C.c_int, c_vp, c_vp, C.c_int, c_vp, C.c_int, c_vp]
        L.plg_world_destroy.argtypes = [c_vp]
        L.plg_last_error.restype = C.c_char_p
        L.plg_rrtc_solve.argtypes = [c_vp, C.c_int, c_vp, c_vp, c_vp, c_vp, C.c_double, C.c_char_p, C.c_double,
                                     C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.POINTER(PlgResult), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
        L.dropin_set_nn_path.argtypes = [C.c_int]
        L.plg_set_two_phase.argtypes = [C.c_uint32, C.c_int]
        L.dropin_rrtc_bench.argtypes = [c_vp, C.c_int, c_vp, c_vp, C.c_uint32, c_vp, c_vp, C.c_double, C.c_char_p,
                                        C.c_double, C.c_uint64, C.c_uint32, C.POINTER(C.c_double),
                                        C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        _dropin = L
    return _dropin


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(ROOT, "pr-ompl_b200", "libprgpu_plugin.so"))
        L.plg_world_create.restype = c_vp
        L.plg_world_create.argtypes = [C.c_int, C.c_int, c_vp, C.c_int, c_vp, c_vp, C.c_int, c_vp, C.c_int, c_vp]
        L.plg_world_destroy.argtypes = [c_vp]
        L.plg_last_error.restype = C.c_char_p
        L.plg_rrtc_solve.argtypes = [c_vp, C.c_int, c_vp, c_vp, c_vp, c_vp, C.c_double, C.c_char_p, C.c_double,
                                     C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.POINTER(PlgResult), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
        L.plg_set_two_phase.argtypes = [C.c_uint32, C.c_int]
        L.plg_batched_rrtc_solve.argtypes = [c_vp, C.c_int, c_vp, c_vp, C.c_uint32, c_vp, c_vp, C.c_double,
                                             C.c_char_p, C.c_double, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                             c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_uint32]
        L.plg_selftest_interfaces.argtypes = [c_vp, c_vp, C.c_uint32, c_vp, c_vp, c_vp]
        L.plg_nnf_run.argtypes = [c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, C.c_uint32, c_vp,
                                  C.c_uint32, C.c_double, C.POINTER(PlgNnfResult), c_vp, c_vp, C.c_uint32, c_vp]
        L.plg_nnf_run_host.argtypes = [c_vp, C.c_int, C.c_int, C.POINTER(C.c_int)] + L.plg_nnf_run.argtypes[1:]
        _lib = L
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


class PluginWorld:
    def __init__(self, dh, link, xyzr, sph, box, device=0, library=None):
        """library: None = libprgpu_plugin.so (the re-implemented planner classes); dropin() = the reference's
        own planner source over the GPU interfaces."""
        self.L = library if library is not None else lib()
        self.keep = [np.ascontiguousarray(dh, dtype=np.float64), np.ascontiguousarray(link, dtype=np.int32),
                     np.ascontiguousarray(xyzr, dtype=np.float64), np.ascontiguousarray(sph, dtype=np.float64),
                     np.ascontiguousarray(box, dtype=np.float64)]
        dh, link, xyzr, sph, box = self.keep
        self.dof = dh.shape[0]
        self.h = self.L.plg_world_create(device, self.dof, _p(dh), len(link), _p(link), _p(xyzr), len(sph), _p(sph),
                                        len(box), _p(box))
        if not self.h:
            raise RuntimeError(self.L.plg_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None) and getattr(self, "L", None) is not None:
            self.L.plg_world_destroy(self.h)
            self.h = None

    def rrtc_solve(self, start, goal, lo, hi, rng=0.0, extension_type="ext-con", resolution_fraction=0.01, seed=0,
                   query=0, max_iters=1000, max_nodes=20000, max_path=4096):
        dim = self.dof
        start = np.ascontiguousarray(start, dtype=np.float64)
        goal = np.ascontiguousarray(goal, dtype=np.float64)
        lo = np.ascontiguousarray(np.broadcast_to(lo, (dim,)), dtype=np.float64)
        hi = np.ascontiguousarray(np.broadcast_to(hi, (dim,)), dtype=np.float64)
        res = PlgResult()
        ss, gs = np.zeros((max_nodes, dim)), np.zeros((max_nodes, dim))
        sp, gp = np.zeros(max_nodes, dtype=np.int32), np.zeros(max_nodes, dtype=np.int32)
        path = np.zeros((max_path, dim))
        pd = np.zeros(4, dtype=np.uint32)
        rc = self.L.plg_rrtc_solve(self.h, dim, _p(lo), _p(hi), _p(start), _p(goal), float(rng),
                                  extension_type.encode(), float(resolution_fraction), int(seed), int(query),
                                  int(max_iters), int(max_nodes), int(max_path), C.byref(res), _p(ss), _p(sp),
                                  _p(gs), _p(gp), _p(path), _p(pd))
        if rc != 0:
            raise RuntimeError(self.L.plg_last_error().decode())
        assert res.overflow == 0
        return dict(status=res.status, iterations=res.iterations, start_states=ss[:res.n_start].copy(),
                    start_parent=sp[:res.n_start].copy(), goal_states=gs[:res.n_goal].copy(),
                    goal_parent=gp[:res.n_goal].copy(), path=path[:res.path_len].copy()), pd

    def selftest(self, states):
        states = np.ascontiguousarray(states, dtype=np.float64)
        n = states.shape[0]
        v = np.zeros(n, dtype=np.uint8)
        m = np.zeros(max(n - 1, 1), dtype=np.uint8)
        nn = np.zeros(n, dtype=np.uint32)
        rc = self.L.plg_selftest_interfaces(self.h, _p(states), n, _p(v), _p(m), _p(nn))
        if rc != 0:
            raise RuntimeError(self.L.plg_last_error().decode())
        return v, m[:n - 1], nn

    def nnf_run(self, W, M, k, D, seed, ref_pose12, ik_states, solve_seconds=20.0, cap_path=65536):
        ref = np.ascontiguousarray(ref_pose12, dtype=np.float64).reshape(-1, 12)
        ik = np.ascontiguousarray(ik_states, dtype=np.float64).reshape(-1, self.dof)
        R = (W - 1) * (D + 1) + 1
        res = PlgNnfResult()
        counts = np.zeros(R, dtype=np.uint32)
        pv = np.zeros(cap_path, dtype=np.uint32)
        ps = np.zeros((cap_path, self.dof))
        rc = self.L.plg_nnf_run(self.h, W, M, k, D, seed, _p(ref), len(ref), _p(ik), len(ik), float(solve_seconds),
                               C.byref(res), _p(counts), _p(pv), cap_path, _p(ps))
        if rc != 0:
            raise RuntimeError(self.L.plg_last_error().decode())
        return dict(status=res.status, lazy_iters=res.lazy_iters, n_evaluated=res.n_evaluated,
                    n_collisions=res.n_collisions, final_error=res.final_error, goal_cost=res.goal_cost,
                    layer_counts=counts, path=pv[:res.path_len].copy(), path_states=ps[:res.path_len].copy())

    def nnf_run_host(self, W, M, k, D, seed, ref_pose12, ik_states, metric=0, with_world=False, solve_seconds=20.0,
                     cap_path=65536):
        """NNFrechet::NNFrechet in the form setup() chooses: metric 1 (rotation term) or with_world=False
        must end in the host-callback form."""
        ref = np.ascontiguousarray(ref_pose12, dtype=np.float64).reshape(-1, 12)
        ik = np.ascontiguousarray(ik_states, dtype=np.float64).reshape(-1, self.dof)
        R = (W - 1) * (D + 1) + 1
        res = PlgNnfResult()
        counts = np.zeros(R, dtype=np.uint32)
        pv = np.zeros(cap_path, dtype=np.uint32)
        ps = np.zeros((cap_path, self.dof))
        host = C.c_int(-1)
        rc = self.L.plg_nnf_run_host(self.h, int(metric), int(with_world), C.byref(host), W, M, k, D, seed, _p(ref),
                                     len(ref), _p(ik), len(ik), float(solve_seconds), C.byref(res), _p(counts),
                                     _p(pv), cap_path, _p(ps))
        if rc != 0:
            raise RuntimeError(self.L.plg_last_error().decode())
        return dict(status=res.status, lazy_iters=res.lazy_iters, n_evaluated=res.n_evaluated,
                    n_collisions=res.n_collisions, final_error=res.final_error, goal_cost=res.goal_cost,
                    layer_counts=counts, path=pv[:res.path_len].copy(), path_states=ps[:res.path_len].copy(),
                    host_form=host.value)

    def set_two_phase(self, first_iters, clear_between=False):
        self.L.plg_set_two_phase(int(first_iters), int(clear_between))

    def batched_rrtc_solve(self, starts, goals, lo, hi, rng=0.0, extension_type="ext-con", resolution_fraction=0.01,
                           seed=0, query_base=0, max_iters=1000, max_nodes=2048, max_path=512):
        dim = self.dof
        starts = np.ascontiguousarray(starts, dtype=np.float64)
        goals = np.ascontiguousarray(goals, dtype=np.float64)
        n = len(starts)
        lo = np.ascontiguousarray(np.broadcast_to(lo, (dim,)), dtype=np.float64)
        hi = np.ascontiguousarray(np.broadcast_to(hi, (dim,)), dtype=np.float64)
        st = np.zeros(n, dtype=np.int32)
        it, ns, ng, pl = (np.zeros(n, dtype=np.uint32) for _ in range(4))
        paths = np.zeros((n, max_path, dim))
        rc = self.L.plg_batched_rrtc_solve(self.h, dim, _p(lo), _p(hi), n, _p(starts), _p(goals), float(rng),
                                           extension_type.encode(), float(resolution_fraction), int(seed),
                                           int(query_base), int(max_iters), int(max_nodes), _p(st), _p(it), _p(ns),
                                           _p(ng), _p(pl), _p(paths), max_path)
        if rc != 0:
            raise RuntimeError(self.L.plg_last_error().decode())
        assert (pl <= max_path).all()
        return dict(status=st, iterations=it, n_start=ns, n_goal=ng, paths=[paths[i, :pl[i]].copy() for i in range(n)])

    def rrtc_bench(self, starts, goals, lo, hi, rng=0.0, extension_type="ext-con", resolution_fraction=0.01, seed=0,
                   max_iters=1000):
        """Drop-in A only: plans the queries one after the other; returns (seconds, solved, iterations)."""
        dim = self.dof
        starts = np.ascontiguousarray(starts, dtype=np.float64)
        goals = np.ascontiguousarray(goals, dtype=np.float64)
        lo = np.ascontiguousarray(np.broadcast_to(lo, (dim,)), dtype=np.float64)
        hi = np.ascontiguousarray(np.broadcast_to(hi, (dim,)), dtype=np.float64)
        sec, solved, iters = C.c_double(), C.c_uint32(), C.c_uint64()
        rc = self.L.dropin_rrtc_bench(self.h, dim, _p(lo), _p(hi), len(starts), _p(starts), _p(goals), float(rng),
                                      extension_type.encode(), float(resolution_fraction), int(seed), int(max_iters),
                                      C.byref(sec), C.byref(solved), C.byref(iters))
        if rc != 0:
            raise RuntimeError(self.L.plg_last_error().decode())
        return sec.value, solved.value, iters.value
