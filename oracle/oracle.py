"""TEST INFRASTRUCTURE ONLY: ctypes wrapper around oracle/libpdg_oracle.so (pdg_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (paralleldg_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODEL_UNIFORM, MODEL_GGM, MODEL_LOGLIN = 0, 1, 2
PRIOR_MBC, PRIOR_EDGEPENALTY, PRIOR_JUMPPENALTY, PRIOR_UNIFORM = 0, 1, 2, 3
KIND_PARALLEL, KIND_SINGLE = 0, 1


class CModel(C.Structure):
    _fields_ = [("p", C.c_int32), ("model", C.c_int32), ("prior", C.c_int32), ("pad0", C.c_int32),
                ("prior_a", C.c_double), ("prior_b", C.c_double),
                ("A", C.c_void_p), ("Dm", C.c_void_p), ("n", C.c_int64), ("delta", C.c_double),
                ("data", C.c_void_p), ("levels", C.c_void_p), ("nrows", C.c_int64),
                ("cell_alpha", C.c_double)]


class CRec(C.Structure):
    _fields_ = [("i", C.c_int32), ("k", C.c_int32), ("type", C.c_int32), ("node", C.c_int32),
                ("U", C.c_int32), ("Uadj", C.c_int32)]


REC_DTYPE = np.dtype([("i", "<i4"), ("k", "<i4"), ("type", "<i4"), ("node", "<i4"),
                      ("U", "<i4"), ("Uadj", "<i4")])


class CReplay(C.Structure):
    _fields_ = [("it_node", C.c_void_p), ("it_mtype_raw", C.c_void_p), ("it_aux", C.c_void_p),
                ("it_off", C.c_void_p), ("mv_U", C.c_void_p), ("mv_Uadj", C.c_void_p),
                ("mv_u", C.c_void_p)]


class CMovesOut(C.Structure):
    _fields_ = [("cap", C.c_int64), ("count", C.c_int64),
                ("iter", C.c_void_p), ("U", C.c_void_p), ("Uadj", C.c_void_p), ("acc", C.c_void_p),
                ("dlik", C.c_void_p), ("dprior", C.c_void_p), ("logq", C.c_void_p), ("u", C.c_void_p)]


def build(force=False):
    so = os.path.join(_HERE, "libpdg_oracle.so")
    src = os.path.join(_HERE, "pdg_oracle.c")
    hdr = os.path.join(_HERE, "pdg_oracle.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpdg_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.pdgo_score.restype = C.c_double
        L.pdgo_score.argtypes = [C.c_void_p, C.c_void_p]
        L.pdgo_prior_partial.restype = C.c_double
        L.pdgo_prior_partial.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.pdgo_chain_create.restype = C.c_void_p
        L.pdgo_chain_create.argtypes = [C.c_void_p, C.c_int]
        L.pdgo_chain_destroy.argtypes = [C.c_void_p]
        L.pdgo_chain_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdgo_chain_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdgo_chain_set_cliques.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.pdgo_chain_init_logl.restype = C.c_double
        L.pdgo_chain_init_logl.argtypes = [C.c_void_p, C.c_int]
        L.pdgo_chain_randomize.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]
        L.pdgo_chain_run.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_uint64,
                                     C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int64, C.c_void_p, C.c_void_p]
        L.pdgo_run_many.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_int64, C.c_int64,
                                    C.c_uint64, C.c_int] + [C.c_void_p] * 5 + [C.c_int64] + [C.c_void_p] * 2
        L.pdgo_tree_from_forest.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdgo_philox4x32_10.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox(c0, c1, c2, c3, k0, k1):
    out = np.zeros(4, np.uint32)
    lib().pdgo_philox4x32_10(c0, c1, c2, c3, k0, k1, _ptr(out))
    return out


_PRIORS = {"mbc": PRIOR_MBC, "edgepenalty": PRIOR_EDGEPENALTY, "jumppenalty": PRIOR_JUMPPENALTY,
           "uniform": PRIOR_UNIFORM}


def prior_spec(graph_prior):
    """mh_parallel.py:317-332 get_prior -> (kind, a, b)."""
    name = str(graph_prior[0]).lower()
    defaults = {"mbc": [2.0, 4.0], "edgepenalty": [0.001], "jumppenalty": [0.25], "uniform": []}
    if name not in defaults:
        name = "mbc"
    dv = defaults[name]
    params = list(graph_prior[1:]) if len(graph_prior) > 1 else dv
    params = [float(x) for x in params[:len(dv)]]
    params += [0.0] * (2 - len(params))
    return _PRIORS[name], params[0], params[1]


class Model(object):
    """Holds the numpy buffers and the C struct of one model (keeps them alive)."""

    def __init__(self, p, model, prior=("mbc", 2.0, 4.0)):
        self.p = int(p)
        self.W = (self.p + 63) // 64
        self.c = CModel()
        self.c.p = self.p
        self.c.model = model
        pk, pa, pb = prior_spec(list(prior))
        self.c.prior, self.c.prior_a, self.c.prior_b = pk, pa, pb
        self._keep = []

    @classmethod
    def uniform(cls, p, prior=("uniform",)):
        return cls(p, MODEL_UNIFORM, prior)

    @classmethod
    def ggm(cls, X, D=None, delta=1.0, prior=("mbc", 2.0, 4.0)):
        """seqdist:334-343: SS = X^T X (uncentred); A = D + SS (ggm:36)."""
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        n, p = X.shape
        if D is None:
            D = np.identity(p)
        D = np.ascontiguousarray(np.asarray(D, dtype=np.float64))
        SS = X.T @ X
        A = np.ascontiguousarray(D + SS)
        m = cls(p, MODEL_GGM, prior)
        m.A, m.D, m.n, m.delta = A, D, n, float(delta)
        m.c.A, m.c.Dm, m.c.n, m.c.delta = _ptr(A).value, _ptr(D).value, n, float(delta)
        return m

    @classmethod
    def loglin(cls, data, levels, cell_alpha=1.0, prior=("mbc", 2.0, 4.0)):
        data = np.asarray(data)
        n, p = data.shape
        col = np.ascontiguousarray(data.T.astype(np.uint8))
        lv = np.ascontiguousarray(np.asarray(levels, dtype=np.int32))
        m = cls(p, MODEL_LOGLIN, prior)
        m.data, m.levels, m.n, m.cell_alpha = col, lv, n, float(cell_alpha)
        m.c.data, m.c.levels, m.c.nrows, m.c.cell_alpha = _ptr(col).value, _ptr(lv).value, n, float(cell_alpha)
        return m

    def ref(self):
        return C.byref(self.c)

    def bits(self, s):
        out = np.zeros(self.W, np.uint64)
        for g in s:
            out[int(g) >> 6] |= np.uint64(1) << np.uint64(int(g) & 63)
        return out

    def score(self, s):
        b = s if isinstance(s, np.ndarray) else self.bits(s)
        return lib().pdgo_score(self.ref(), _ptr(np.ascontiguousarray(b)))

    def prior_partial(self, csize, ssize, extra):
        return lib().pdgo_prior_partial(self.ref(), csize, ssize, extra)


class Chain(object):
    def __init__(self, model, use_cache=True):
        self.model = model
        self.h = lib().pdgo_chain_create(model.ref(), 1 if use_cache else 0)
        self.p, self.W = model.p, model.W

    def __del__(self):
        try:
            if self.h:
                lib().pdgo_chain_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_state(self, edges, bits):
        e = np.ascontiguousarray(edges, dtype=np.int32).reshape(self.p - 1, 2)
        b = np.ascontiguousarray(bits, dtype=np.uint64).reshape(self.p, self.W)
        return lib().pdgo_chain_set_state(self.h, _ptr(e), _ptr(b))

    def get_state(self):
        e = np.zeros((self.p - 1, 2), np.int32)
        b = np.zeros((self.p, self.W), np.uint64)
        lib().pdgo_chain_get_state(self.h, _ptr(e), _ptr(b))
        return e, b

    def set_cliques(self, bits):
        b = np.ascontiguousarray(bits, dtype=np.uint64).reshape(-1, self.W)
        return lib().pdgo_chain_set_cliques(self.h, _ptr(b), b.shape[0])

    def set_singletons(self):
        b = np.zeros((self.p, self.W), np.uint64)
        for v in range(self.p):
            b[v, v >> 6] = np.uint64(1) << np.uint64(v & 63)
        return self.set_cliques(b)

    def init_logl(self, kind=KIND_PARALLEL):
        return lib().pdgo_chain_init_logl(self.h, kind)

    def randomize(self, seed, chain_id, it):
        return lib().pdgo_chain_randomize(self.h, seed, chain_id, it)

    def run(self, kind, it_begin, it_end, logl, k, replay=None, seed=0, chain_id=0,
            rec_cap=0, want_moves=0):
        """returns dict(rc, k, recs, rec_bits, moves)"""
        recs = np.zeros(rec_cap, REC_DTYPE) if rec_cap else None
        rbits = np.zeros((rec_cap, 2, self.W), np.uint64) if rec_cap else None
        nrec = C.c_int64(0)
        kk = C.c_int64(k)
        rp = None
        keep = []
        if replay is not None:
            rp = CReplay()
            for f, dt in (("it_node", np.int32), ("it_mtype_raw", np.int32), ("it_aux", np.int32),
                          ("it_off", np.int64), ("mv_U", np.int32), ("mv_Uadj", np.int32),
                          ("mv_u", np.float64)):
                a = np.ascontiguousarray(replay[f], dtype=dt)
                keep.append(a)
                setattr(rp, f, _ptr(a).value if a.size else None)
        mv = None
        mva = {}
        if want_moves:
            mv = CMovesOut()
            mv.cap, mv.count = want_moves, 0
            for f, dt in (("iter", np.int32), ("U", np.int32), ("Uadj", np.int32), ("acc", np.int32),
                          ("dlik", np.float64), ("dprior", np.float64), ("logq", np.float64),
                          ("u", np.float64)):
                mva[f] = np.zeros(want_moves, dt)
                setattr(mv, f, _ptr(mva[f]).value)
        rc = lib().pdgo_chain_run(self.h, kind, it_begin, it_end,
                                  C.byref(rp) if rp is not None else None, seed, chain_id,
                                  _ptr(logl), C.byref(kk), _ptr(recs), _ptr(rbits), rec_cap,
                                  C.byref(nrec), C.byref(mv) if mv is not None else None)
        out = dict(rc=rc, k=kk.value, n_rec=nrec.value)
        if rec_cap:
            out["recs"] = recs[:min(nrec.value, rec_cap)]
            out["rec_bits"] = rbits[:min(nrec.value, rec_cap)]
        if want_moves:
            n = min(mv.count, want_moves)
            out["moves"] = {f: a[:n] for f, a in mva.items()}
            out["n_moves"] = mv.count
        return out


def run_replay(model, log, kind=None, want_moves=True):
    """Replays a reference log (oracle/ref_shims.trace_chain) through the C oracle.
    Returns dict(rc, logl, k, recs, rec_bits, moves)."""
    M = int(log["n_samples"])
    R = int(log["randomize"])
    kind = int(log["single"]) if kind is None else kind
    ch = Chain(model)
    states = {int(i): j for j, i in enumerate(log["state_iter"])}
    ch.set_state(log["state_edges"][0], log["state_bits"][0])
    logl = np.zeros(M, np.float64)
    logl[0] = ch.init_logl(kind)
    k = 1 if kind == KIND_SINGLE else 0
    nmv = len(log["mv_U"])
    all_recs, all_bits, all_moves = [], [], []
    it = 1
    rc = 0
    while it < M and rc == 0:
        if it % R == 0:
            j = states[it]
            ch.set_state(log["state_edges"][j], log["state_bits"][j])
        end = min((it // R + 1) * R, M)
        out = ch.run(kind, it, end, logl, k, replay=log, rec_cap=(end - it) * (model.p + 1),
                     want_moves=(end - it) * (2 * model.p + 2) if want_moves else 0)
        rc = out["rc"]
        k = out["k"]
        all_recs.append(out["recs"])
        all_bits.append(out["rec_bits"])
        if want_moves:
            all_moves.append(out["moves"])
        it = end
    res = dict(rc=rc, logl=logl, k=k, recs=np.concatenate(all_recs), rec_bits=np.concatenate(all_bits),
               final_state=ch.get_state())
    if want_moves:
        res["moves"] = {f: np.concatenate([m[f] for m in all_moves]) for f in all_moves[0]}
    return res


def run_many(model, kind, n_chains, n_samples, randomize, seed, chain0=0, n_threads=1,
             want_logl=False, rec_cap=0, want_state=False):
    """Philox production mode over many chains (the CPU baseline)."""
    n_prop = np.zeros(n_chains, np.int64)
    n_acc = np.zeros(n_chains, np.int64)
    logl = np.zeros((n_chains, n_samples), np.float64) if want_logl else None
    recs = np.zeros((n_chains, rec_cap), REC_DTYPE) if rec_cap else None
    rbits = np.zeros((n_chains, rec_cap, 2, model.W), np.uint64) if rec_cap else None
    fe = np.zeros((n_chains, model.p - 1, 2), np.int32) if want_state else None
    fb = np.zeros((n_chains, model.p, model.W), np.uint64) if want_state else None
    rc = lib().pdgo_run_many(model.ref(), kind, n_chains, chain0, n_samples, randomize, seed, n_threads,
                             _ptr(n_prop), _ptr(n_acc), _ptr(logl), _ptr(recs), _ptr(rbits), rec_cap,
                             _ptr(fe), _ptr(fb))
    return dict(rc=rc, n_prop=n_prop, n_acc=n_acc, logl=logl, recs=recs, rec_bits=rbits,
                final_edges=fe, final_bits=fb)
