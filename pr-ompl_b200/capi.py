"""ctypes binding of include/prgpu.h (libprgpu.so).  No CPU fallback: if the library is missing
or no sm_100 device is present, calls raise PrgpuError."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
INVALID_INDEX = 0xFFFFFFFF
MODE_OMPL_DISCRETE = 0
MODE_NNF_ALPHA = 1

c_dp = C.POINTER(C.c_double)
c_u32p = C.POINTER(C.c_uint32)
c_i32p = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)
c_vp = C.c_void_p


class PrgpuError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "libprgpu.so")


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libprgpu.so (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-j8", "-C", os.path.join(_HERE, "csrc")], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    return lib_path()


_lib = None

# name -> (restype, argtypes); every symbol include/prgpu.h declares
SIGNATURES = {
    "prgpu_last_error": (C.c_char_p, []),
    "prgpu_version": (C.c_int, []),
    "prgpu_device_count": (C.c_int, []),
    "prgpu_knn_create": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.POINTER(c_vp)]),
    "prgpu_knn_destroy": (C.c_int, [c_vp]),
    "prgpu_knn_clear": (C.c_int, [c_vp]),
    "prgpu_knn_add": (C.c_int, [c_vp, c_vp, C.c_uint64]),
    "prgpu_knn_add_device": (C.c_int, [c_vp, c_vp, C.c_uint64, c_vp]),
    "prgpu_knn_size": (C.c_int, [c_vp, C.POINTER(C.c_uint64)]),
    "prgpu_knn_set_index_base": (C.c_int, [c_vp, C.c_uint32]),
    "prgpu_knn_get_points": (C.c_int, [c_vp, C.c_uint64, C.c_uint64, c_vp]),
    "prgpu_knn_nearest_k": (C.c_int, [c_vp, c_vp, C.c_uint64, C.c_int, c_vp, c_vp, c_vp]),
    "prgpu_knn_nearest_k_device": (C.c_int, [c_vp, c_vp, C.c_uint64, C.c_int, c_vp, c_vp, c_vp, c_vp]),
    "prgpu_knn_set_search_mode": (C.c_int, [c_vp, C.c_int]),
    "prgpu_knn_build_index": (C.c_int, [c_vp]),
    "prgpu_knn_index_stats": (C.c_int, [c_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_float), c_u32p,
                                        C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "prgpu_knn_search_stats": (C.c_int, [c_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "prgpu_knn_set_profiling": (C.c_int, [c_vp, C.c_int]),
    "prgpu_knn_last_kernel_ms": (C.c_int, [c_vp, C.POINTER(C.c_float)]),
    "prgpu_world_set_profiling": (C.c_int, [c_vp, C.c_int]),
    "prgpu_world_last_kernel_ms": (C.c_int, [c_vp, C.POINTER(C.c_float)]),
    "prgpu_topk_merge_device": (C.c_int, [C.c_int, c_vp, c_vp, C.c_int, C.c_uint64, C.c_int, c_vp, c_vp,
                                          c_vp]),
    "prgpu_world_create": (C.c_int, [C.c_int, C.c_int, c_vp, C.c_int, c_vp, c_vp, C.c_int, c_vp, C.c_int,
                                     c_vp, C.POINTER(c_vp)]),
    "prgpu_world_destroy": (C.c_int, [c_vp]),
    "prgpu_world_dof": (C.c_int, [c_vp]),
    "prgpu_check_motion_batch": (C.c_int, [c_vp, c_vp, c_vp, C.c_uint64, C.c_double, C.c_int, C.c_int,
                                           c_vp]),
    "prgpu_check_motion_batch_device": (C.c_int, [c_vp, c_vp, c_vp, C.c_uint64, C.c_double, C.c_int,
                                                  C.c_int, c_vp, c_vp]),
    "prgpu_is_valid_batch": (C.c_int, [c_vp, c_vp, C.c_uint64, c_vp]),
    "prgpu_is_valid_batch_device": (C.c_int, [c_vp, c_vp, C.c_uint64, c_vp, c_vp]),
    "prgpu_fk_batch": (C.c_int, [c_vp, c_vp, C.c_uint64, c_vp]),
    "prgpu_fk_batch_device": (C.c_int, [c_vp, c_vp, C.c_uint64, c_vp, c_vp]),
    "prgpu_rrtc_batch_create": (C.c_int, [c_vp, c_vp, C.c_uint32, C.POINTER(c_vp)]),
    "prgpu_rrtc_batch_destroy": (C.c_int, [c_vp]),
    "prgpu_rrtc_batch_run": (C.c_int, [c_vp, c_vp, c_vp]),
    "prgpu_rrtc_batch_run_device": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "prgpu_rrtc_batch_summary": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "prgpu_rrtc_batch_fetch": (C.c_int, [c_vp, C.c_uint32, c_vp, c_vp, c_vp, c_vp, C.POINTER(C.c_uint32),
                                         c_vp, C.c_uint32]),
    "prgpu_comm_unique_id": (C.c_int, [c_vp]),
    "prgpu_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_int, c_vp, C.POINTER(c_vp)]),
    "prgpu_comm_init_custom": (C.c_int, [C.c_int, C.c_int, C.c_int, c_vp, c_vp, C.POINTER(c_vp)]),
    "prgpu_comm_destroy": (C.c_int, [c_vp]),
    "prgpu_comm_rank": (C.c_int, [c_vp]),
    "prgpu_comm_size": (C.c_int, [c_vp]),
    "prgpu_comm_allgather": (C.c_int, [c_vp, c_vp, c_vp, C.c_uint64, c_vp]),
    "prgpu_comm_enable_peer_windows": (C.c_int, [c_vp, C.c_uint64]),
    "prgpu_comm_stats2": (C.c_int, [c_vp, C.POINTER(C.c_uint64)]),
    "prgpu_comm_stats": (C.c_int, [c_vp, C.POINTER(C.c_uint64)]),
    "prgpu_knn_nearest_k_sharded_device": (C.c_int, [c_vp, c_vp, c_vp, C.c_uint64, C.c_int, c_vp, c_vp, c_vp]),
    "prgpu_knn_nearest_k_query_sharded_device": (C.c_int, [c_vp, c_vp, c_vp, C.c_uint64, C.c_int, c_vp, c_vp, c_vp]),
    "prgpu_knn_nearest_k_query_sharded": (C.c_int, [c_vp, c_vp, c_vp, C.c_uint64, C.c_int, c_vp, c_vp]),
    "prgpu_knn_nearest_k_sharded": (C.c_int, [c_vp, c_vp, c_vp, C.c_uint64, C.c_int, c_vp, c_vp]),
    "prgpu_check_motion_batch_sharded_device": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_uint64, C.c_double, C.c_int,
                                                          C.c_int, c_vp, c_vp]),
    "prgpu_diag_launch_count": (C.c_uint64, []),
    "prgpu_diag_mma_tf32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "prgpu_nnf_create": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp, C.POINTER(c_vp)]),
    "prgpu_nnf_destroy": (C.c_int, [c_vp]),
    "prgpu_nnf_graph_info": (C.c_int, [c_vp, c_u32p, c_u32p, c_u32p]),
    "prgpu_nnf_get_knn": (C.c_int, [c_vp, c_vp]),
    "prgpu_nnf_get_vertices": (C.c_int, [c_vp, c_vp, c_vp]),
    "prgpu_nnf_knn_stats": (C.c_int, [c_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "prgpu_nnf_get_edges": (C.c_int, [c_vp, c_vp, c_vp]),
    "prgpu_nnf_solve": (C.c_int, [c_vp, C.c_uint32, c_vp, c_vp, C.c_uint32, c_vp]),
    "prgpu_diag_knn_scan": (C.c_int, [c_vp, c_vp, C.c_uint32, c_vp, C.c_uint32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "prgpu_diag_fk_centre_error": (C.c_int, [c_vp, c_vp, C.c_uint64, c_vp, C.POINTER(C.c_double)]),
    "prgpu_diag_kd_leaf_capacity": (C.c_int, [C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "prgpu_nnf_create_graph": (C.c_int, [C.c_int, C.c_int, c_vp, C.c_int, c_vp, c_vp, C.POINTER(c_vp)]),
    "prgpu_nnf_set_weights": (C.c_int, [c_vp, c_vp]),
    "prgpu_nnf_set_edge_evaluator": (C.c_int, [c_vp, c_vp, c_vp]),
    "prgpu_nnf_set_mode": (C.c_int, [c_vp, C.c_int]),
    "prgpu_nnf_band_info": (C.c_int, [c_vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), c_u32p]),
    "prgpu_nnf_loop_stats": (C.c_int, [c_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), c_vp]),
    "prgpu_nnf_last_dp_ms": (C.c_int, [c_vp, C.POINTER(C.c_float)]),
}


class NnfParams(C.Structure):
    _fields_ = [("num_nn", C.c_int32), ("discretization", C.c_int32), ("check_resolution", C.c_double)]


class NnfResult(C.Structure):
    _fields_ = [("n_vertices", C.c_uint32), ("n_edges", C.c_uint32), ("n_ik", C.c_uint32),
                ("lazy_iters", C.c_uint32), ("solved", C.c_int32), ("final_error", C.c_double),
                ("goal_cost", C.c_double), ("path_len", C.c_uint32), ("n_evaluated", C.c_uint32),
                ("n_collisions", C.c_uint32), ("overflow", C.c_uint32)]


class RrtcParams(C.Structure):
    _fields_ = [("dim", C.c_int32), ("lo", C.c_double * 8), ("hi", C.c_double * 8), ("range", C.c_double),
                ("resolution_fraction", C.c_double), ("ext_first", C.c_int32), ("ext_second", C.c_int32),
                ("seed", C.c_uint64), ("query_id_base", C.c_uint64), ("max_iters", C.c_uint32),
                ("max_nodes", C.c_uint32)]


def lib():
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise PrgpuError("libprgpu.so is not built (run __graft_entry__.build()); there is no CPU "
                             "fallback for the prgpu hot path")
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise PrgpuError("prgpu error %d: %s" % (rc, lib().prgpu_last_error().decode()))


def _ptr(a):
    """host numpy array, torch tensor (device or host) or raw int address -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())


class KnnIndex:
    """Tree configuration buffer + nearest / nearestK (ompl::NearestNeighbors semantics)."""

    def __init__(self, dim, device=0, capacity=0):
        self.dim = dim
        self.device = device
        h = c_vp()
        _check(lib().prgpu_knn_create(device, dim, capacity, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.prgpu_knn_destroy(self.h)
        self.h = None

    __del__ = close

    def clear(self):
        _check(lib().prgpu_knn_clear(self.h))

    def size(self):
        n = C.c_uint64()
        _check(lib().prgpu_knn_size(self.h, C.byref(n)))
        return n.value

    def build_index(self):
        """k-d index over the current nodes, now (otherwise built once the brute-force searches have cost as much)."""
        _check(lib().prgpu_knn_build_index(self.h))

    def diag_scan(self, queries, thresh, cap):
        """prgpu_diag_knn_scan: the production scan's D, the fp32 re-cut keys and the error bounds."""
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, self.dim)
        nq = len(q)
        t = np.ascontiguousarray(np.broadcast_to(thresh, (nq,)), dtype=np.float32).copy()
        counts = np.zeros(nq, dtype=np.uint32)
        idx = np.zeros((nq, cap), dtype=np.uint32)
        D = np.zeros((nq, cap), dtype=np.float32)
        k32 = np.zeros((nq, cap), dtype=np.float32)
        eps = np.zeros((nq, 2))
        _check(lib().prgpu_diag_knn_scan(self.h, _ptr(q), nq, _ptr(t), cap, _ptr(counts), _ptr(idx), _ptr(D),
                                         _ptr(k32), _ptr(eps)))
        return dict(thresh=t, counts=counts, idx=idx, D=D, key32=k32, eps_tc=eps[:, 0], eps_32=eps[:, 1])

    def set_search_mode(self, force_exact):
        """0 auto (k-d index, tensor-core filter or exact kernel, whichever is cheapest), 1 / True exact fp64 kernel only,
        2 brute force without the index, 3 the index whenever it can answer.  Results are identical in every mode."""
        _check(lib().prgpu_knn_set_search_mode(self.h, int(force_exact)))

    def search_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(lib().prgpu_knn_search_stats(self.h, C.byref(a), C.byref(b)))
        return dict(filter_calls=a.value, fallback_queries=b.value)

    def index_stats(self):
        n, ms, b, c, v = C.c_uint64(), C.c_float(), C.c_uint32(), C.c_uint64(), C.c_uint64()
        _check(lib().prgpu_knn_index_stats(self.h, C.byref(n), C.byref(ms), C.byref(b), C.byref(c), C.byref(v)))
        return dict(indexed_nodes=n.value, build_ms=ms.value, builds=b.value, index_calls=c.value,
                    leaves_visited=v.value)

    def set_profiling(self, on=True):
        _check(lib().prgpu_knn_set_profiling(self.h, int(on)))

    def last_kernel_ms(self):
        ms = C.c_float()
        _check(lib().prgpu_knn_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def set_index_base(self, base):
        _check(lib().prgpu_knn_set_index_base(self.h, base))

    def add(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, self.dim)
        _check(lib().prgpu_knn_add(self.h, _ptr(pts), pts.shape[0]))

    def add_device(self, pts_dev, n, stream=0):
        _check(lib().prgpu_knn_add_device(self.h, _ptr(pts_dev), n, C.c_void_p(stream)))

    def points(self, first=0, n=None):
        n = self.size() - first if n is None else n
        out = np.empty((n, self.dim), dtype=np.float64)
        _check(lib().prgpu_knn_get_points(self.h, first, n, _ptr(out)))
        return out

    def nearest_k(self, queries, k, lo=None, want_dist=True):
        queries = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, self.dim)
        nq = queries.shape[0]
        idx = np.empty((nq, k), dtype=np.uint32)
        dist = np.empty((nq, k), dtype=np.float64) if want_dist else None
        if lo is not None:
            lo = np.ascontiguousarray(lo, dtype=np.uint32)
        _check(lib().prgpu_knn_nearest_k(self.h, _ptr(queries), nq, k, _ptr(lo), _ptr(idx), _ptr(dist)))
        return (idx, dist) if want_dist else idx

    def nearest_k_sharded_device(self, comm, q_dev, nq, k, idx_dev, dist_dev, stream=0):
        """Collective: this handle holds one shard of the tree (set_index_base = first global index)."""
        _check(lib().prgpu_knn_nearest_k_sharded_device(self.h, comm.h, _ptr(q_dev), nq, k, _ptr(idx_dev),
                                                        _ptr(dist_dev), C.c_void_p(stream)))

    def nearest_k_query_sharded_device(self, comm, q_dev, nq_local, k, idx_all_dev, dist_all_dev, stream=0):
        """Collective: every rank holds the whole tree and searches its own nq_local queries; all ranks receive
        all answers ([nranks][nq_local][k])."""
        _check(lib().prgpu_knn_nearest_k_query_sharded_device(self.h, comm.h, _ptr(q_dev), nq_local, k,
                                                              _ptr(idx_all_dev), _ptr(dist_all_dev), C.c_void_p(stream)))

    def nearest_k_query_sharded(self, comm, queries, k):
        queries = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, self.dim)
        nr = comm.nranks
        idx = np.empty((nr, len(queries), k), dtype=np.uint32)
        dist = np.empty((nr, len(queries), k), dtype=np.float64)
        _check(lib().prgpu_knn_nearest_k_query_sharded(self.h, comm.h, _ptr(queries), len(queries), k, _ptr(idx),
                                                       _ptr(dist)))
        return idx, dist

    def nearest_k_sharded(self, comm, queries, k):
        queries = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, self.dim)
        idx = np.empty((len(queries), k), dtype=np.uint32)
        dist = np.empty((len(queries), k), dtype=np.float64)
        _check(lib().prgpu_knn_nearest_k_sharded(self.h, comm.h, _ptr(queries), len(queries), k, _ptr(idx),
                                                 _ptr(dist)))
        return idx, dist

    def nearest_k_device(self, q_dev, nq, k, idx_dev, dist_dev=None, lo_dev=None, stream=0):
        _check(lib().prgpu_knn_nearest_k_device(self.h, _ptr(q_dev), nq, k, _ptr(lo_dev), _ptr(idx_dev),
                                                _ptr(dist_dev), C.c_void_p(stream)))


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, c_vp, c_vp, c_vp, C.c_uint64, c_vp)


class Comm:
    """prgpu_comm_*: the communicator behind the sharded entry points (NCCL bound at run time, or a
    caller-supplied all-gather)."""

    def __init__(self, device, rank, nranks, unique_id=None, allgather=None):
        h = c_vp()
        self._cb = None
        if allgather is not None:
            self._cb = ALLGATHER_FN(allgather)      # keep the trampoline alive
            _check(lib().prgpu_comm_init_custom(device, rank, nranks, C.cast(self._cb, c_vp), None, C.byref(h)))
        else:
            buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
            _check(lib().prgpu_comm_init(device, rank, nranks, C.cast(buf, c_vp), C.byref(h)))
        self.h, self.rank, self.nranks = h, rank, nranks

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        _check(lib().prgpu_comm_unique_id(C.cast(buf, c_vp)))
        return bytes(buf)

    def allgather(self, send, recv, bytes_per_rank, stream=0):
        _check(lib().prgpu_comm_allgather(self.h, _ptr(send), _ptr(recv), bytes_per_rank, C.c_void_p(stream)))

    def enable_peer_windows(self, nbytes):
        """Collective.  True if every rank mapped every window (the fused search + exchange is then used)."""
        return lib().prgpu_comm_enable_peer_windows(self.h, int(nbytes)) == 0

    def fused_steps(self):
        v = C.c_uint64()
        _check(lib().prgpu_comm_stats2(self.h, C.byref(v)))
        return v.value

    def redo_steps(self):
        v = C.c_uint64()
        _check(lib().prgpu_comm_stats(self.h, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.prgpu_comm_destroy(self.h)
        self.h = None

    __del__ = close


def topk_merge_device(device, idx_parts, dist_parts, nparts, nq, k, idx_out, dist_out, stream=0):
    _check(lib().prgpu_topk_merge_device(device, _ptr(idx_parts), _ptr(dist_parts), nparts, nq, k,
                                         _ptr(idx_out), _ptr(dist_out), C.c_void_p(stream)))


class World:
    """Synthetic world (DH arm with link spheres vs sphere/box obstacles) on the device."""

    def __init__(self, dh, sphere_link, sphere_xyzr, obs_spheres, obs_boxes, device=0):
        dh = np.ascontiguousarray(dh, dtype=np.float64)
        link = np.ascontiguousarray(sphere_link, dtype=np.int32)
        xyzr = np.ascontiguousarray(sphere_xyzr, dtype=np.float64)
        os_ = np.ascontiguousarray(obs_spheres, dtype=np.float64).reshape(-1, 4)
        ob_ = np.ascontiguousarray(obs_boxes, dtype=np.float64).reshape(-1, 6)
        self.dof = dh.shape[0]
        self.device = device
        h = c_vp()
        _check(lib().prgpu_world_create(device, self.dof, _ptr(dh), len(link), _ptr(link), _ptr(xyzr),
                                        len(os_), _ptr(os_), len(ob_), _ptr(ob_), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.prgpu_world_destroy(self.h)
        self.h = None

    __del__ = close

    def set_profiling(self, on=True):
        _check(lib().prgpu_world_set_profiling(self.h, int(on)))

    def last_kernel_ms(self):
        ms = C.c_float()
        _check(lib().prgpu_world_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def diag_fk_centre_error(self, states):
        """(max over link spheres of |fp32-chain centre - fp64-chain centre| per state, delta)"""
        q = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, self.dof)
        err = np.zeros(len(q))
        delta = C.c_double()
        _check(lib().prgpu_diag_fk_centre_error(self.h, _ptr(q), len(q), _ptr(err), C.byref(delta)))
        return err, delta.value

    def check_motion_batch(self, a, b, resolution, mode=MODE_OMPL_DISCRETE, check_first=False):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, self.dof)
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1, self.dof)
        valid = np.empty(a.shape[0], dtype=np.uint8)
        _check(lib().prgpu_check_motion_batch(self.h, _ptr(a), _ptr(b), a.shape[0], resolution, mode,
                                              int(check_first), _ptr(valid)))
        return valid

    def check_motion_batch_device(self, a_dev, b_dev, n, resolution, valid_dev, mode=MODE_OMPL_DISCRETE,
                                  check_first=False, stream=0):
        _check(lib().prgpu_check_motion_batch_device(self.h, _ptr(a_dev), _ptr(b_dev), n, resolution, mode,
                                                     int(check_first), _ptr(valid_dev), C.c_void_p(stream)))

    def check_motion_batch_sharded_device(self, comm, a_dev, b_dev, n, resolution, valid_dev,
                                          mode=MODE_OMPL_DISCRETE, check_first=False, stream=0):
        """Collective: the batch is replicated, every rank validates its slice and receives all n verdicts."""
        _check(lib().prgpu_check_motion_batch_sharded_device(self.h, comm.h, _ptr(a_dev), _ptr(b_dev), n,
                                                             float(resolution), mode, int(check_first),
                                                             _ptr(valid_dev), C.c_void_p(stream)))

    def is_valid_batch(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.dof)
        valid = np.empty(q.shape[0], dtype=np.uint8)
        _check(lib().prgpu_is_valid_batch(self.h, _ptr(q), q.shape[0], _ptr(valid)))
        return valid

    def fk_batch(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.dof)
        pose = np.empty((q.shape[0], 3, 4), dtype=np.float64)
        _check(lib().prgpu_fk_batch(self.h, _ptr(q), q.shape[0], _ptr(pose)))
        return pose


EXT_TYPES = {"ext-ext": (0, 0), "ext-con": (0, 1), "con-ext": (1, 0), "con-con": (1, 1)}


class RrtcBatch:
    """Many independent pr_ompl::RRTConnect queries against one world (one thread block each)."""

    def __init__(self, world, n_queries, lo, hi, rng=0.0, extension_type="ext-con", resolution_fraction=0.01,
                 seed=0, query_id_base=0, max_iters=1000, max_nodes=2048):
        if extension_type not in EXT_TYPES:
            raise ValueError("Invalid extension type string.")  # RRTConnect::setExtensionTypesString
        self.world = world
        self.n = n_queries
        self.dim = world.dof
        self.max_nodes = max_nodes
        p = RrtcParams()
        p.dim = self.dim
        lo = np.broadcast_to(np.asarray(lo, dtype=np.float64), (self.dim,))
        hi = np.broadcast_to(np.asarray(hi, dtype=np.float64), (self.dim,))
        for i in range(self.dim):
            p.lo[i] = lo[i]
            p.hi[i] = hi[i]
        p.range = rng
        p.resolution_fraction = resolution_fraction
        p.ext_first, p.ext_second = EXT_TYPES[extension_type]
        p.seed, p.query_id_base = seed, query_id_base
        p.max_iters, p.max_nodes = max_iters, max_nodes
        h = c_vp()
        _check(lib().prgpu_rrtc_batch_create(world.h, C.byref(p), n_queries, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.prgpu_rrtc_batch_destroy(self.h)
        self.h = None

    __del__ = close

    def run(self, starts, goals):
        starts = np.ascontiguousarray(starts, dtype=np.float64).reshape(self.n, self.dim)
        goals = np.ascontiguousarray(goals, dtype=np.float64).reshape(self.n, self.dim)
        _check(lib().prgpu_rrtc_batch_run(self.h, _ptr(starts), _ptr(goals)))

    def run_device(self, starts_dev, goals_dev, stream=0):
        _check(lib().prgpu_rrtc_batch_run_device(self.h, _ptr(starts_dev), _ptr(goals_dev), C.c_void_p(stream)))

    def summary(self):
        st = np.empty(self.n, dtype=np.int32)
        it = np.empty(self.n, dtype=np.uint32)
        ns = np.empty(self.n, dtype=np.uint32)
        ng = np.empty(self.n, dtype=np.uint32)
        _check(lib().prgpu_rrtc_batch_summary(self.h, _ptr(st), _ptr(it), _ptr(ns), _ptr(ng)))
        return dict(status=st, iterations=it, n_start=ns, n_goal=ng)

    def fetch(self, query, ns, ng, max_path=8192):
        ss = np.zeros((max(ns, 1), self.dim))
        sp = np.zeros(max(ns, 1), dtype=np.int32)
        gs = np.zeros((max(ng, 1), self.dim))
        gp = np.zeros(max(ng, 1), dtype=np.int32)
        path = np.zeros((max_path, self.dim))
        plen = C.c_uint32()
        _check(lib().prgpu_rrtc_batch_fetch(self.h, query, _ptr(ss), _ptr(sp), _ptr(gs), _ptr(gp),
                                            C.byref(plen), _ptr(path), max_path))
        return dict(start_states=ss[:ns], start_parent=sp[:ns], goal_states=gs[:ng], goal_parent=gp[:ng],
                    path=path[:plen.value].copy())


class Nnf:
    """NNF planner state on the device: NN graph + tensor-product DAG (setup) and LazySP (solve)."""

    def __init__(self, world, ref_pose12, layer_counts, ik_states, num_nn=5, discretization=3,
                 check_resolution=0.05):
        self.world = world
        self.k = num_nn
        ref = np.ascontiguousarray(ref_pose12, dtype=np.float64).reshape(-1, 12)
        counts = np.ascontiguousarray(layer_counts, dtype=np.uint32)
        ik = np.ascontiguousarray(ik_states, dtype=np.float64).reshape(-1, world.dof)
        assert len(counts) == len(ref) and counts.sum() == len(ik)
        p = NnfParams(num_nn, discretization, check_resolution)
        h = c_vp()
        _check(lib().prgpu_nnf_create(world.h, C.byref(p), len(ref), _ptr(ref), _ptr(counts), _ptr(ik),
                                      C.byref(h)))
        self.h = h
        nv, ne, ni = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().prgpu_nnf_graph_info(self.h, C.byref(nv), C.byref(ne), C.byref(ni)))
        self.n_vertices, self.n_edges, self.n_ik = nv.value, ne.value, ni.value

    @classmethod
    def host_form(cls, dof, layer_counts, ik_states, num_nn=5, discretization=3, check_resolution=0.05, device=0):
        """Host-callback form (prgpu_nnf_create_graph): the caller supplies node weights (set_weights) and
        edge verdicts (set_edge_evaluator); the device does findKnnNodes, the NN graph and the searches."""
        self = cls.__new__(cls)
        self.world = None
        self.dof = dof
        self.k = num_nn
        counts = np.ascontiguousarray(layer_counts, dtype=np.uint32)
        ik = np.ascontiguousarray(ik_states, dtype=np.float64).reshape(-1, dof)
        assert counts.sum() == len(ik)
        self.R = len(counts)
        p = NnfParams(num_nn, discretization, check_resolution)
        h = c_vp()
        _check(lib().prgpu_nnf_create_graph(device, dof, C.byref(p), len(counts), _ptr(counts), _ptr(ik), C.byref(h)))
        self.h = h
        nv, ne, ni = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().prgpu_nnf_graph_info(self.h, C.byref(nv), C.byref(ne), C.byref(ni)))
        self.n_vertices, self.n_edges, self.n_ik = nv.value, ne.value, ni.value
        return self

    def states(self):
        st = np.empty((self.n_vertices, self.dof if self.world is None else self.world.dof))
        _check(lib().prgpu_nnf_get_vertices(self.h, _ptr(st), None))
        return st

    def set_weights(self, weights):
        """weights[v, r] = task distance between reference pose r and the pose of NN-graph vertex v."""
        w = np.ascontiguousarray(weights, dtype=np.float64)
        assert w.shape == (self.n_vertices, self.R)
        _check(lib().prgpu_nnf_set_weights(self.h, _ptr(w)))

    def set_edge_evaluator(self, fn):
        """fn(from[n, dof], to[n, dof]) -> array of n booleans, True = the edge is collision-free."""
        dof = self.dof if self.world is None else self.world.dof
        proto = C.CFUNCTYPE(C.c_int, c_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint32,
                            C.POINTER(C.c_uint8))

        def thunk(_user, a, b, n, out):
            try:
                fa = np.ctypeslib.as_array(a, shape=(n, dof))
                fb = np.ctypeslib.as_array(b, shape=(n, dof))
                ok = np.asarray(fn(fa.copy(), fb.copy())).astype(np.uint8)
                for i in range(n):
                    out[i] = int(ok[i])
                return 0
            except Exception:  # noqa: BLE001 -- reported through the C ABI's error return
                return 1

        self._eval_cb = proto(thunk)  # keep alive
        _check(lib().prgpu_nnf_set_edge_evaluator(self.h, C.cast(self._eval_cb, c_vp), None))

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.prgpu_nnf_destroy(self.h)
        self.h = None

    __del__ = close

    def knn(self):
        out = np.empty((self.n_ik, self.k), dtype=np.uint32)
        _check(lib().prgpu_nnf_get_knn(self.h, _ptr(out)))
        return out

    def vertices(self):
        st = np.empty((self.n_vertices, self.world.dof))
        po = np.empty((self.n_vertices, 12))
        _check(lib().prgpu_nnf_get_vertices(self.h, _ptr(st), _ptr(po)))
        return st, po

    def knn_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(lib().prgpu_nnf_knn_stats(self.h, C.byref(a), C.byref(b)))
        return dict(filter_calls=a.value, fallback_queries=b.value)

    def edges(self):
        """NN-graph edge list (creation order) and flags (bit 0 evaluated, bit 1 knocked out)."""
        ed = np.empty((self.n_edges, 2), dtype=np.uint32)
        fl = np.empty(self.n_edges, dtype=np.uint8)
        _check(lib().prgpu_nnf_get_edges(self.h, _ptr(ed), _ptr(fl)))
        return ed, fl

    def solve(self, max_lazy_iters=1000, cap_path=65536):
        res = NnfResult()
        path = np.zeros(cap_path, dtype=np.uint32)
        trace = np.zeros(max_lazy_iters, dtype=np.float64)
        _check(lib().prgpu_nnf_solve(self.h, max_lazy_iters, C.byref(res), _ptr(path), cap_path, _ptr(trace)))
        out = {f[0]: getattr(res, f[0]) for f in NnfResult._fields_}
        out["path"] = path[:res.path_len].copy()
        out["goal_cost_trace"] = trace[:res.lazy_iters].copy()
        return out

    def set_mode(self, mode):
        """0 = banded column-wise DP (default), 1 = full-table anti-diagonal wavefront."""
        _check(lib().prgpu_nnf_set_mode(self.h, mode))

    def band_info(self):
        tau, cells, reb = C.c_double(), C.c_uint64(), C.c_uint32()
        _check(lib().prgpu_nnf_band_info(self.h, C.byref(tau), C.byref(cells), C.byref(reb)))
        return dict(tau=tau.value, cells=cells.value, rebuilds=reb.value)

    def loop_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        cyc = np.zeros(6, dtype=np.uint64)
        _check(lib().prgpu_nnf_loop_stats(self.h, C.byref(a), C.byref(b), _ptr(cyc)))
        names = ("hop_chase", "per_hop_pass", "edge_evaluation", "bookkeeping", "repair")
        return dict(launches=a.value, repaired_vertices=b.value, phase_cycles={n: int(c) for n, c in zip(names, cyc)},
                    wide_waves_handed_to_full_sweep=int(cyc[5]))

    def last_dp_ms(self):
        ms = C.c_float()
        _check(lib().prgpu_nnf_last_dp_ms(self.h, C.byref(ms)))
        return ms.value
