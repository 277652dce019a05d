"""ctypes binding of libpdg_b200.so (include/pdg_b200.h).  There is NO CPU fallback: when the
library or a CUDA device is missing every entry point raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (PDG_LIB selects a developer build variant, e.g. one made by `python -m paralleldg_b200.build --variant nb2 ...`)
LIB_PATH = os.environ.get("PDG_LIB") or os.path.join(_HERE, "libpdg_b200.so")

MODEL_UNIFORM, MODEL_GGM, MODEL_LOGLIN = 0, 1, 2
PRIOR_MBC, PRIOR_EDGEPENALTY, PRIOR_JUMPPENALTY, PRIOR_UNIFORM = 0, 1, 2, 3
KIND_PARALLEL, KIND_SINGLE = 0, 1

UPDATE_DTYPE = np.dtype([("i", "<i4"), ("k", "<i4"), ("node", "<u2"), ("U", "<u2"), ("Uadj", "<i2"),
                         ("type", "u1"), ("pad", "u1")])
assert UPDATE_DTYPE.itemsize == 16

EXPORTS = [
    "pdg_abi_version", "pdg_create", "pdg_destroy", "pdg_last_error", "pdg_set_stream", "pdg_set_prior",
    "pdg_set_model_uniform", "pdg_set_model_ggm", "pdg_set_model_loglin", "pdg_set_state", "pdg_get_state",
    "pdg_set_cliques", "pdg_randomize", "pdg_init_logl", "pdg_run", "pdg_sample", "pdg_sync",
    "pdg_get_counts", "pdg_get_logl", "pdg_compact_updates", "pdg_get_updates", "pdg_enable_move_log",
    "pdg_get_move_log", "pdg_score_sets", "pdg_edge_tallies", "pdg_launch_count",
    "pdg_enable_timing", "pdg_get_timing", "pdg_get_work", "pdg_set_seed", "pdg_host_alloc", "pdg_host_free", "pdg_probe", "pdg_collect",
    "pdg_size_series", "pdg_edge_ess",
]


E_ARG, E_CUDA, E_STATE, E_REPLAY, E_OVERFLOW, E_NUMERIC, E_NOMEM = -1, -2, -3, -4, -5, -6, -7


class NativeError(RuntimeError):
    """carries the pdg_error code of include/pdg_b200.h in `.code` (0 when not from the library)"""

    def __init__(self, msg, code=0):
        RuntimeError.__init__(self, msg)
        self.code = int(code)


class CReplay(C.Structure):
    _fields_ = [("n_samples", C.c_int64), ("it_node", C.c_void_p), ("it_mtype", C.c_void_p),
                ("it_aux", C.c_void_p), ("it_off", C.c_void_p), ("mv_U", C.c_void_p),
                ("mv_Uadj", C.c_void_p), ("mv_u", C.c_void_p), ("n_moves_total", C.c_int64)]


_LIB = None


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise NativeError("libpdg_b200.so is not built (run `python -m paralleldg_b200.build`); "
                          "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, u32, u64, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.c_uint64, C.c_double
    sig = {
        "pdg_abi_version": (C.c_int, []),
        "pdg_create": (C.c_int, [C.POINTER(vp), i32, i32, i64, u32, u64, i64, i64]),
        "pdg_destroy": (C.c_int, [vp]),
        "pdg_last_error": (C.c_char_p, [vp]),
        "pdg_set_stream": (C.c_int, [vp, vp]),
        "pdg_set_prior": (C.c_int, [vp, i32, dbl, dbl]),
        "pdg_set_model_uniform": (C.c_int, [vp]),
        "pdg_set_model_ggm": (C.c_int, [vp, vp, vp, i64, dbl, vp]),
        "pdg_set_model_loglin": (C.c_int, [vp, vp, vp, i64, dbl, i32, vp, vp, vp]),
        "pdg_set_state": (C.c_int, [vp, i64, i64, vp, vp]),
        "pdg_get_state": (C.c_int, [vp, i64, i64, vp, vp]),
        "pdg_set_cliques": (C.c_int, [vp, vp, i32]),
        "pdg_randomize": (C.c_int, [vp, u32]),
        "pdg_init_logl": (C.c_int, [vp, i32]),
        "pdg_run": (C.c_int, [vp, i32, i64, i64, vp]),
        "pdg_sample": (C.c_int, [vp, i32, i64]),
        "pdg_sync": (C.c_int, [vp]),
        "pdg_get_counts": (C.c_int, [vp, vp, vp]),
        "pdg_get_logl": (C.c_int, [vp, vp]),
        "pdg_compact_updates": (C.c_int, [vp, C.POINTER(i64), vp]),
        "pdg_get_updates": (C.c_int, [vp, vp, vp]),
        "pdg_enable_move_log": (C.c_int, [vp, i64]),
        "pdg_get_move_log": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
        "pdg_score_sets": (C.c_int, [vp, vp, i64, vp]),
        "pdg_edge_tallies": (C.c_int, [vp, i64, vp]),
        "pdg_launch_count": (i64, [vp]),
        "pdg_enable_timing": (C.c_int, [vp, i32]),
        "pdg_get_timing": (C.c_int, [vp, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(i64), C.POINTER(i64)]),
        "pdg_get_work": (C.c_int, [vp, C.POINTER(i64)]),
        "pdg_set_seed": (C.c_int, [vp, u64, u32]),
        "pdg_host_alloc": (C.c_int, [C.POINTER(vp), i64]),
        "pdg_host_free": (C.c_int, [vp]),
        "pdg_probe": (C.c_int, [i32, i32, C.POINTER(dbl)]),
        "pdg_collect": (C.c_int, [vp, C.POINTER(i64), vp, vp, vp, vp, i64, vp]),
        "pdg_size_series": (C.c_int, [vp, vp]),
        "pdg_edge_ess": (C.c_int, [vp, i64, dbl, dbl, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


class PinnedPool(object):
    """Caching allocator of page-locked host memory for the result arrays (the same idea as
    torch's CachingHostAllocator, without torch): numpy arrays handed out here return their block
    to the pool when they are garbage collected, so back-to-back runs reuse the pinned pages."""

    def __init__(self):
        self.free = {}          # rounded size -> [ptr]

    @staticmethod
    def _round(nbytes):
        return max(4096, 1 << (int(nbytes) - 1).bit_length()) if nbytes > (1 << 20) else 1 << 20

    def empty(self, shape, dtype):
        import weakref
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        if nbytes == 0:
            return np.empty(shape, dtype)
        size = self._round(nbytes)
        lst = self.free.get(size)
        if lst:
            ptr = lst.pop()
        else:
            p = C.c_void_p()
            if lib().pdg_host_alloc(C.byref(p), size) != 0:
                return np.empty(shape, dtype)          # pinning failed: pageable memory still works
            ptr = p.value
        buf = (C.c_char * nbytes).from_address(ptr)
        weakref.finalize(buf, self._release, size, ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def _release(self, size, ptr):
        self.free.setdefault(size, []).append(ptr)


PINNED = PinnedPool()


def _p(a):
    """host numpy array / torch tensor / None -> void*"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError("expected numpy array or torch tensor")


class Context(object):
    """One GPU, many chains.  Thin, stateful mirror of the C-ABI."""

    def __init__(self, p, n_chains, n_samples, seed=0, chain0=0, device=0, rec_cap=None):
        self.L = lib()
        self.p, self.W = int(p), (int(p) + 63) // 64
        self.n_chains, self.n_samples = int(n_chains), int(n_samples)
        self.rec_cap = int(rec_cap if rec_cap is not None else 2 * n_samples)
        self.h = C.c_void_p()
        rc = self.L.pdg_create(C.byref(self.h), int(device), self.p, self.n_chains, int(chain0),
                               int(seed) & 0xFFFFFFFFFFFFFFFF, self.n_samples, self.rec_cap)
        if rc != 0:
            self.h = C.c_void_p()
            raise NativeError("pdg_create failed with code %d (no CUDA device? out of memory?)" % rc, rc)
        self.mvlog_cap = 0

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.pdg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self.L.pdg_last_error(self.h)
            raise NativeError("libpdg_b200 error %d: %s" % (rc, msg.decode() if msg else ""), rc)

    def set_seed(self, seed, chain0=0):
        self._ck(self.L.pdg_set_seed(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF, int(chain0)))

    def set_stream(self, stream_ptr):
        self._ck(self.L.pdg_set_stream(self.h, C.c_void_p(stream_ptr)))

    def set_prior(self, kind, a=0.0, b=0.0):
        self._ck(self.L.pdg_set_prior(self.h, int(kind), float(a), float(b)))

    def set_model_uniform(self):
        self._ck(self.L.pdg_set_model_uniform(self.h))

    def set_model_ggm(self, A, D, n, delta, gconst):
        self._keep = (A, D, gconst)
        self._ck(self.L.pdg_set_model_ggm(self.h, _p(A), _p(D), int(n), float(delta), _p(gconst)))

    def set_model_loglin(self, data_colmajor, levels, nrows, cell_alpha, ktab=None, alphatab=None, lgtab=None):
        self._keep = (data_colmajor, levels, ktab, alphatab, lgtab)
        nk = 0 if ktab is None else int(len(ktab))
        self._ck(self.L.pdg_set_model_loglin(self.h, _p(data_colmajor), _p(levels), int(nrows), float(cell_alpha),
                                             nk, _p(ktab), _p(alphatab), _p(lgtab)))

    def set_state(self, edges, bits, chain_begin=0):
        e = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, self.p - 1, 2)
        b = np.ascontiguousarray(bits, dtype=np.uint64).reshape(-1, self.p, self.W)
        assert e.shape[0] == b.shape[0]
        self._ck(self.L.pdg_set_state(self.h, int(chain_begin), e.shape[0], _p(e), _p(b)))

    def get_state(self, chain_begin=0, count=None):
        count = self.n_chains - chain_begin if count is None else count
        e = np.zeros((count, self.p - 1, 2), np.int32)
        b = np.zeros((count, self.p, self.W), np.uint64)
        self._ck(self.L.pdg_get_state(self.h, int(chain_begin), int(count), _p(e), _p(b)))
        return e, b

    def set_cliques(self, bits=None):
        if bits is None:
            self._ck(self.L.pdg_set_cliques(self.h, None, 0))
        else:
            b = np.ascontiguousarray(bits, dtype=np.uint64).reshape(-1, self.W)
            self._ck(self.L.pdg_set_cliques(self.h, _p(b), b.shape[0]))

    def randomize(self, it):
        self._ck(self.L.pdg_randomize(self.h, int(it)))

    def init_logl(self, kind):
        self._ck(self.L.pdg_init_logl(self.h, int(kind)))

    def run(self, kind, it_begin, it_end, replay=None):
        rp = None
        if replay is not None:
            rp = CReplay()
            rp.n_samples = int(replay["n_samples"])
            keep = []
            for f, dt in (("it_node", np.int32), ("it_mtype", np.int32), ("it_aux", np.int32),
                          ("it_off", np.int64), ("mv_U", np.int32), ("mv_Uadj", np.int32), ("mv_u", np.float64)):
                a = np.ascontiguousarray(replay[f], dtype=dt)
                keep.append(a)
                setattr(rp, f, a.ctypes.data if a.size else None)
            rp.n_moves_total = int(len(keep[4]))
            self._rp_keep = keep
        self._ck(self.L.pdg_run(self.h, int(kind), int(it_begin), int(it_end), C.byref(rp) if rp is not None else None))

    def sample(self, kind, randomize):
        self._ck(self.L.pdg_sample(self.h, int(kind), int(randomize)))

    def sync(self):
        self._ck(self.L.pdg_sync(self.h))

    def counts(self):
        a = np.zeros(self.n_chains, np.int64)
        b = np.zeros(self.n_chains, np.int64)
        self._ck(self.L.pdg_get_counts(self.h, _p(a), _p(b)))
        return a, b

    def logl(self, out=None):
        out = PINNED.empty((self.n_chains, self.n_samples), np.float64) if out is None else out
        self._ck(self.L.pdg_get_logl(self.h, _p(out)))
        return out

    def updates(self):
        """-> (offsets[n_chains+1], records[total], bits[total, 2, W])"""
        tot = C.c_int64(0)
        off = np.zeros(self.n_chains + 1, np.int64)
        self._ck(self.L.pdg_compact_updates(self.h, C.byref(tot), _p(off)))
        rec = PINNED.empty((tot.value,), UPDATE_DTYPE)
        bits = PINNED.empty((tot.value, 2, self.W), np.uint64)
        self._ck(self.L.pdg_get_updates(self.h, _p(rec), _p(bits)))
        return off, rec, bits

    def collect(self, want_logl=True, want_updates=True, want_tallies=False, burnin=0):
        """everything a finished run returns, with overlapped copies and one wait (pdg_collect):
        -> dict(n_proposals, n_accepted, offsets, logl, records, bits, tallies)"""
        tot = C.c_int64(0)
        out = dict(n_proposals=np.zeros(self.n_chains, np.int64), n_accepted=np.zeros(self.n_chains, np.int64),
                   offsets=np.zeros(self.n_chains + 1, np.int64), logl=None, records=None, bits=None, tallies=None)
        if want_logl:
            out["logl"] = PINNED.empty((self.n_chains, self.n_samples), np.float64)
        if want_tallies:
            out["tallies"] = PINNED.empty((self.p, self.p), np.int64)
        self._ck(self.L.pdg_collect(self.h, C.byref(tot), _p(out["n_proposals"]), _p(out["n_accepted"]), _p(out["offsets"]),
                                    _p(out["logl"]), int(burnin), _p(out["tallies"])))
        if want_updates:
            out["records"] = PINNED.empty((tot.value,), UPDATE_DTYPE)
            out["bits"] = PINNED.empty((tot.value, 2, self.W), np.uint64)
        self._ck(self.L.pdg_get_updates(self.h, _p(out["records"]), _p(out["bits"])))
        return out

    def enable_move_log(self, cap):
        self._ck(self.L.pdg_enable_move_log(self.h, int(cap)))
        self.mvlog_cap = int(cap)

    def move_log(self):
        n = (self.n_chains, self.mvlog_cap)
        out = dict(dlik=np.zeros(n), dprior=np.zeros(n), logq=np.zeros(n), u=np.zeros(n),
                   acc=np.zeros(n, np.int32), U=np.zeros(n, np.int32))
        self._ck(self.L.pdg_get_move_log(self.h, _p(out["dlik"]), _p(out["dprior"]), _p(out["logq"]),
                                         _p(out["u"]), _p(out["acc"]), _p(out["U"])))
        return out

    def score_sets(self, bits):
        b = np.ascontiguousarray(bits, dtype=np.uint64).reshape(-1, self.W)
        out = np.zeros(b.shape[0], np.float64)
        self._ck(self.L.pdg_score_sets(self.h, _p(b), b.shape[0], _p(out)))
        return out

    def edge_tallies(self, burnin=0):
        out = np.zeros((self.p, self.p), np.int64)
        self._ck(self.L.pdg_edge_tallies(self.h, int(burnin), _p(out)))
        return out

    def size_series(self):
        """number of edges at every list position of every chain (trajectory.py:202-211), int32 [chains, n_samples]"""
        out = np.empty((self.n_chains, self.n_samples), np.int32)
        self._ck(self.L.pdg_size_series(self.h, _p(out)))
        return out

    def edge_ess(self, burnin=0, lo=0.05, hi=0.95):
        """per-chain median edge-indicator ESS and the number of edges it is taken over"""
        ess = np.empty(self.n_chains, np.float64)
        ne = np.empty(self.n_chains, np.int32)
        self._ck(self.L.pdg_edge_ess(self.h, int(burnin), float(lo), float(hi), _p(ess), _p(ne)))
        return ess, ne

    def enable_timing(self, on=True):
        self._ck(self.L.pdg_enable_timing(self.h, 1 if on else 0))

    def timing(self):
        a, b, na, nb = C.c_double(0), C.c_double(0), C.c_int64(0), C.c_int64(0)
        self._ck(self.L.pdg_get_timing(self.h, C.byref(a), C.byref(b), C.byref(na), C.byref(nb)))
        return dict(ms_run=a.value, ms_randomize=b.value, n_run=na.value, n_randomize=nb.value)

    def work(self):
        w = (C.c_int64 * 16)()
        self._ck(self.L.pdg_get_work(self.h, w))
        return dict(bytes=int(w[0]), flops=int(w[1]) / 6.0, sets=int(w[2]), sum_k=int(w[3]),
                    sets_requested=int(w[4]), sum_k_requested=int(w[5]), sets_sparse=int(w[6]), sets_derived=int(w[7]), phase=[int(x) for x in w[8:16]])

    def launch_count(self):
        return int(self.L.pdg_launch_count(self.h))


PROBES = {"fp64_gflops": 0, "l2_read_gbs": 1, "hbm_read_gbs": 2, "red_shared_free_gps": 3, "red_shared_random_gps": 4}


def probe(which, device=0):
    """machine ceilings measured by libpdg_b200.so itself (include/pdg_b200.h: pdg_probe)"""
    v = C.c_double(0.0)
    rc = lib().pdg_probe(int(device), PROBES[which], C.byref(v))
    if rc != 0:
        raise NativeError("pdg_probe(%s) failed with code %d" % (which, rc), rc)
    return v.value
