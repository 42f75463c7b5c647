"""
ctypes binding of the C-ABI in ``include/pz.h`` (``libpz_b200.so``, built
in-tree by ``__graft_entry__.build()``).

There is NO CPU fallback: if the shared library is missing, or no CUDA device
is present, every entry point raises.  The reference has no native layer at
all -- the functions below are the calls a maintainer would add behind
``percolate.hpc``.
"""

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PZ_B200_LIB") or os.path.join(_HERE, "libpz_b200.so")   # env: A/B builds

PERM_HOST, PERM_DEVICE, PERM_MT19937, PERM_PHILOX, PERM_FEISTEL, PERM_PHILOX_FY = 0, 1, 2, 3, 4, 5
# rng keyword of the Python API -> perm_mode
RNG_MODES = {'mt19937': PERM_MT19937, 'philox': PERM_PHILOX, 'feistel': PERM_FEISTEL,
             'philox_fy': PERM_PHILOX_FY}
FUSE_MICRO, FUSE_CANON = 1, 2
SEEDS_ON_DEVICE = 0x100
ACC_WORDS = 25
CANON_COLS = 7

# every symbol include/pz.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "pz_last_error", "pz_version", "pz_create", "pz_destroy", "pz_device",
    "pz_stream", "pz_synchronize", "pz_set_graph", "pz_row_bytes",
    "pz_run_rows", "pz_run_fused", "pz_reset_accumulators", "pz_micro_runs",
    "pz_micro_export", "pz_micro_import", "pz_micro_finalize", "pz_micro_arrays", "pz_set_ps",
    "pz_convolve", "pz_canonical_statistics_rows", "pz_canon_export",
    "pz_canon_merge", "pz_canon_last_runs", "pz_canon_last_count", "pz_launch_count",
    "pz_make_perms", "pz_profile", "pz_profile_read", "pz_canon_reset", "pz_timer_start", "pz_timer_stop",
    "pz_comm_unique_id", "pz_comm_init", "pz_comm_destroy", "pz_comm_world", "pz_comm_rank", "pz_allreduce",
]
COMM_ID_BYTES = 128


class NativeError(RuntimeError):
    pass


_lib = None
_lib_lock = threading.Lock()


def load():
    """Load ``libpz_b200.so``; raise loudly when it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                "pypercolate_b200: CUDA extension %s not built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`); "
                "there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        c_void_pp = ctypes.POINTER(ctypes.c_void_p)
        vp = ctypes.c_void_p
        i32, i64, ci = ctypes.c_int32, ctypes.c_int64, ctypes.c_int
        L.pz_last_error.restype = ctypes.c_char_p
        L.pz_last_error.argtypes = []
        L.pz_version.restype = ci
        L.pz_create.argtypes = [ci, c_void_pp]
        L.pz_create.restype = ci
        L.pz_destroy.argtypes = [vp]
        L.pz_destroy.restype = None
        L.pz_device.argtypes = [vp]
        L.pz_stream.argtypes = [vp]
        L.pz_stream.restype = vp
        L.pz_synchronize.argtypes = [vp]
        L.pz_set_graph.argtypes = [vp, i32, i32, vp, vp, vp, ci]
        L.pz_row_bytes.argtypes = [vp]
        L.pz_run_rows.argtypes = [vp, i32, ci, vp, vp, vp]
        L.pz_make_perms.argtypes = [vp, i32, ci, vp, vp, ci]
        L.pz_run_fused.argtypes = [vp, i32, ci, vp, ci]
        L.pz_reset_accumulators.argtypes = [vp]
        L.pz_micro_runs.argtypes = [vp]
        L.pz_micro_runs.restype = i64
        L.pz_micro_export.argtypes = [vp, vp, ci]
        L.pz_micro_import.argtypes = [vp, vp, ci, i64]
        L.pz_micro_finalize.argtypes = [vp, vp, vp]
        L.pz_micro_arrays.argtypes = [vp, ctypes.c_double, ctypes.c_double, ctypes.c_double, vp]
        L.pz_set_ps.argtypes = [vp, i32, i32, vp, vp]
        L.pz_convolve.argtypes = [vp, i32, vp, vp]
        L.pz_canonical_statistics_rows.argtypes = [vp, i32, ci, vp, vp, vp]
        L.pz_profile.argtypes = [vp, ci]
        L.pz_profile_read.argtypes = [vp, vp, vp]
        L.pz_canon_export.argtypes = [vp, ctypes.POINTER(i64), vp, vp]
        L.pz_canon_merge.argtypes = [vp, i64, vp, vp]
        L.pz_canon_last_runs.argtypes = [vp, vp]
        L.pz_canon_last_count.argtypes = [vp]
        L.pz_canon_last_count.restype = i32
        L.pz_canon_reset.argtypes = [vp]
        L.pz_timer_start.argtypes = [vp]
        L.pz_timer_stop.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
        L.pz_launch_count.argtypes = [vp]
        L.pz_launch_count.restype = i64
        L.pz_comm_unique_id.argtypes = [vp]
        L.pz_comm_init.argtypes = [vp, ci, ci, vp]
        L.pz_comm_destroy.argtypes = [vp]
        L.pz_comm_world.argtypes = [vp]
        L.pz_comm_rank.argtypes = [vp]
        L.pz_allreduce.argtypes = [vp]
        _lib = L
        return _lib


def _check(rc):
    if rc != 0:
        raise NativeError("libpz_b200: %s (status %d)" % (
            load().pz_last_error().decode("utf-8", "replace"), rc))


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


class Context(object):
    """One device context (one CUDA stream, its scratch and accumulators)."""

    def __init__(self, device=0):
        L = load()
        h = ctypes.c_void_p()
        _check(L.pz_create(int(device), ctypes.byref(h)))
        self._h = h
        self._L = L
        self.device = int(device)
        self.N = self.M = 0
        self.spanning = False
        self.num_p = 0
        self.pmf_M = -1

    def close(self):
        if getattr(self, "_h", None):
            self._L.pz_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- graph ------------------------------------------------------------
    def set_graph(self, lowered):
        eu = np.ascontiguousarray(lowered.eu, dtype=np.int32)
        ev = np.ascontiguousarray(lowered.ev, dtype=np.int32)
        sm = None
        if lowered.side_mask is not None:
            sm = np.ascontiguousarray(lowered.side_mask, dtype=np.uint8)
        _check(self._L.pz_set_graph(self._h, lowered.num_nodes, lowered.num_edges,
                                    _ptr(eu), _ptr(ev), _ptr(sm),
                                    int(lowered.preconnected)))
        self.N, self.M = lowered.num_nodes, lowered.num_edges
        self.spanning = sm is not None
        self.num_p = 0

    @property
    def row_bytes(self):
        return 53 if self.spanning else 52

    def row_dtype(self):
        fields = [('n', '<u4'), ('edge', '<u4')]
        if self.spanning:
            fields.append(('has_spanning_cluster', '?'))
        fields += [('max_cluster_size', '<u4'), ('moments', '<u8', (5,))]
        return np.dtype(fields)

    def _perm_src(self, R, perm_mode, src):
        if perm_mode == PERM_HOST:
            a = np.ascontiguousarray(src, dtype=np.int32)
            if a.shape != (R, self.M):
                raise ValueError("perms must have shape (runs, num_edges)")
            return a, _ptr(a)
        if perm_mode == PERM_DEVICE or (perm_mode & SEEDS_ON_DEVICE):
            return None, ctypes.c_void_p(int(src))
        a = np.ascontiguousarray(src, dtype=np.uint32)
        if a.shape != (R,):
            raise ValueError("seeds must have shape (runs,)")
        return a, _ptr(a)

    # -- materialised runs --------------------------------------------------
    def run_rows(self, R, perm_mode, src, want_perms=False):
        keep, p = self._perm_src(R, perm_mode, src)
        rows = np.empty((R, self.M + 1), dtype=self.row_dtype())
        perms = np.empty((R, self.M), dtype=np.int32) if want_perms else None
        _check(self._L.pz_run_rows(self._h, R, perm_mode, p, _ptr(rows), _ptr(perms)))
        return (rows, perms) if want_perms else rows

    def make_perms(self, R, perm_mode, seeds, out_device_ptr=None):
        """Bond orders of R runs from seeds (device RNG modes)."""
        a = np.ascontiguousarray(seeds, dtype=np.uint32)
        if out_device_ptr is not None:
            _check(self._L.pz_make_perms(self._h, R, perm_mode, _ptr(a),
                                         ctypes.c_void_p(int(out_device_ptr)), 1))
            return None
        out = np.empty((R, self.M), dtype=np.int32)
        _check(self._L.pz_make_perms(self._h, R, perm_mode, _ptr(a), _ptr(out), 0))
        return out

    # -- fused path -----------------------------------------------------------
    def run_fused(self, R, perm_mode, src, flags):
        keep, p = self._perm_src(R, perm_mode, src)
        _check(self._L.pz_run_fused(self._h, R, perm_mode, p, int(flags)))

    def reset_accumulators(self):
        _check(self._L.pz_reset_accumulators(self._h))

    def synchronize(self):
        _check(self._L.pz_synchronize(self._h))

    @property
    def micro_runs(self):
        return int(self._L.pz_micro_runs(self._h))

    def micro_export(self, device_ptr=None):
        if device_ptr is not None:
            _check(self._L.pz_micro_export(self._h, ctypes.c_void_p(int(device_ptr)), 1))
            return None
        out = np.empty((self.M + 1, ACC_WORDS), dtype=np.uint64)
        _check(self._L.pz_micro_export(self._h, _ptr(out), 0))
        return out

    def micro_import(self, src, runs, is_device=False):
        if is_device:
            _check(self._L.pz_micro_import(self._h, ctypes.c_void_p(int(src)), 1, int(runs)))
        else:
            a = np.ascontiguousarray(src, dtype=np.uint64)
            _check(self._L.pz_micro_import(self._h, _ptr(a), 0, int(runs)))

    def micro_finalize(self):
        mean = np.empty((7, self.M + 1), dtype=np.float64)
        var = np.empty((6, self.M + 1), dtype=np.float64)
        _check(self._L.pz_micro_finalize(self._h, _ptr(mean), _ptr(var)))
        return mean, var

    def micro_arrays(self, t_lo, t_hi, norm=1.0):
        """Per-n means and Student-t intervals, divided by ``norm``, as views of one
        host buffer: ``k[S], max[S], max_ci[S, 2], moments[5, S], moments_ci[5, S, 2]``."""
        S = self.M + 1
        buf = np.empty(19 * S, dtype=np.float64)
        _check(self._L.pz_micro_arrays(self._h, float(t_lo), float(t_hi), float(norm), _ptr(buf)))
        return (buf[:S], buf[S:2 * S], buf[2 * S:4 * S].reshape(S, 2),
                buf[14 * S:].reshape(5, S), buf[4 * S:14 * S].reshape(5, S, 2))

    def micro_finalize_device_only(self):
        _check(self._L.pz_micro_finalize(self._h, None, None))

    # -- canonical ------------------------------------------------------------
    def set_ps(self, ps, want_pmf=False, M=None):
        ps = np.ascontiguousarray(ps, dtype=np.float64).reshape(-1)
        M = self.M if M is None else int(M)
        pmf = np.empty((ps.size, M + 1), dtype=np.float64) if want_pmf else None
        _check(self._L.pz_set_ps(self._h, M, ps.size, _ptr(ps), _ptr(pmf)))
        self.num_p = ps.size
        self.pmf_M = M
        return pmf

    def convolve(self, cols):
        cols = np.ascontiguousarray(cols, dtype=np.float64)
        if cols.ndim != 2 or cols.shape[1] != self.pmf_M + 1:
            raise ValueError("cols must have shape (num_cols, num_edges + 1)")
        out = np.empty((cols.shape[0], self.num_p), dtype=np.float64)
        _check(self._L.pz_convolve(self._h, cols.shape[0], _ptr(cols), _ptr(out)))
        return out

    def canonical_statistics_rows(self, rows, f, spanning=None):
        rows = np.ascontiguousarray(rows)
        f = np.ascontiguousarray(f, dtype=np.float64)
        if spanning is None:
            spanning = 'has_spanning_cluster' in rows.dtype.names
        if rows.dtype.itemsize != (53 if spanning else 52) or f.size != rows.size:
            raise ValueError("rows must be packed microcanonical statistics of one run")
        out = np.empty(CANON_COLS, dtype=np.float64)
        _check(self._L.pz_canonical_statistics_rows(self._h, rows.size - 1, int(bool(spanning)),
                                                    _ptr(rows), _ptr(f), _ptr(out)))
        return out

    PHASES = ("perm", "sweep", "accumulate", "canon_runs", "canon_reduce", "rows", "checkpoints")

    def profile(self, enable=True):
        _check(self._L.pz_profile(self._h, int(bool(enable))))

    def profile_read(self):
        ms = np.zeros(len(self.PHASES), dtype=np.float64)
        cnt = np.zeros(len(self.PHASES), dtype=np.int64)
        _check(self._L.pz_profile_read(self._h, _ptr(ms), _ptr(cnt)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.PHASES)}

    def canon_export(self):
        cnt = ctypes.c_int64()
        mean = np.empty((self.num_p, CANON_COLS), dtype=np.float64)
        m2 = np.empty((self.num_p, CANON_COLS), dtype=np.float64)
        _check(self._L.pz_canon_export(self._h, ctypes.byref(cnt), _ptr(mean), _ptr(m2)))
        return int(cnt.value), mean, m2

    def canon_merge(self, count, mean, m2):
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        m2 = np.ascontiguousarray(m2, dtype=np.float64)
        _check(self._L.pz_canon_merge(self._h, int(count), _ptr(mean), _ptr(m2)))

    def canon_replace(self, count, mean, m2):
        _check(self._L.pz_canon_reset(self._h))
        if count:
            self.canon_merge(count, mean, m2)

    def canon_last_runs(self, R):
        last = int(self._L.pz_canon_last_count(self._h))
        if last != R:
            raise NativeError("canon_last_runs: the last device batch held %d runs, not %d" % (last, R))
        out = np.empty((R, self.num_p, CANON_COLS), dtype=np.float64)
        _check(self._L.pz_canon_last_runs(self._h, _ptr(out)))
        return out

    # -- cross-GPU exchange (NCCL through the C-ABI) -----------------------------
    def comm_init(self, world, rank, comm_id):
        """Join the communicator identified by ``comm_id`` (``comm_unique_id()`` of rank 0,
        ``COMM_ID_BYTES`` bytes).  Collective over the ``world`` ranks."""
        buf = bytes(comm_id)
        if len(buf) != COMM_ID_BYTES:
            raise ValueError("comm_id must be %d bytes" % COMM_ID_BYTES)
        _check(self._L.pz_comm_init(self._h, int(world), int(rank), ctypes.c_char_p(buf)))

    def comm_destroy(self):
        _check(self._L.pz_comm_destroy(self._h))

    @property
    def comm_world(self):
        return int(self._L.pz_comm_world(self._h))

    def allreduce(self):
        """Combine the accumulators of every rank's context (collective)."""
        _check(self._L.pz_allreduce(self._h))

    def timer_start(self):
        _check(self._L.pz_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_double()
        _check(self._L.pz_timer_stop(self._h, ctypes.byref(ms)))
        return float(ms.value)

    @property
    def launch_count(self):
        return int(self._L.pz_launch_count(self._h))


def comm_unique_id():
    """Fresh communicator id (rank 0 calls this and hands the bytes to the other ranks)."""
    buf = ctypes.create_string_buffer(COMM_ID_BYTES)
    _check(load().pz_comm_unique_id(buf))
    return buf.raw


# -- one context per (device, graph) ---------------------------------------------
_ctx_lock = threading.Lock()


def context_for(lowered, device=0):
    """Context with ``lowered`` uploaded; cached on the LoweredGraph object."""
    with _ctx_lock:
        ctx = lowered._handles.get(device)
        if ctx is None:
            ctx = Context(device)
            ctx.set_graph(lowered)
            lowered._handles[device] = ctx
        return ctx
