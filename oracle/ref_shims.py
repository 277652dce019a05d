"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* upstream parallelDG reference from /root/reference under the
minimal shim set of SURVEY.md section 8(c) so that it runs on Python 3.12 / numpy 2 /
networkx 3, and traces one chain into a *replay log* (the golden vectors committed under
tests/golden/).  This file only works inside the build container: /root/reference does not
exist on the GPU box, so nothing that runs there may import it.

Shims (each one restores python-2.7-era behaviour, none changes the algorithm):
  1. seaborn / matplotlib stubs             (graph/graph.py:11, auxiliary_functions.py:8-9)
  2. nx.random_tree -> nx.random_labeled_tree (mh_parallel.py:39,196; removed in networkx 3.4)
  3. wishart.log_norm_constant: .tostring() -> .tobytes()  (wishart.py:32; removed in numpy 2)
  4. python-int tree-node ids: JunctionMap.create_t_and_t2clique rebuilt with int() nodes and
     connect/disconnect cast t_node to int, so isinstance(Uadj, int) behaves as on python 2
     (mh_parallel.py:97,140,248,280; junction_tree.py:60; parallel_moves.py:138)
  5. random.seed(seed) next to np.random.seed(seed)  (parallel_moves.py:185 uses `random`)
"""
import os
import random
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("PDG_REFERENCE_ROOT", "/root/reference")

_loaded = None


def load_reference():
    """Returns a namespace with the reference modules (shimmed)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("seaborn", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                 "matplotlib.colors", "matplotlib.cm"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].use = lambda *a, **k: None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import networkx as nx
    if not hasattr(nx, "random_tree"):
        nx.random_tree = lambda n, seed=None: nx.random_labeled_tree(n, seed=seed)

    import scipy.special as scp
    import parallelDG.distributions.wishart as wish

    def log_norm_constant(D, delta, cache={}):
        # wishart.py:17-37 verbatim except .tostring() -> .tobytes()
        K = 0.0
        p = len(D)
        t = (delta + p - 1.0) / 2.0
        K = - t * p * np.log(np.pi)
        if (t, p) not in cache:
            cache[(t, p)] = scp.multigammaln(t, p)
        K += cache[(t, p)]
        tup = np.array(D).ravel().tobytes()
        if tup not in cache:
            (sign, logdet) = np.linalg.slogdet(D)
            cache[tup] = logdet
        K -= t * cache[tup]
        return K
    wish.log_norm_constant = log_norm_constant

    import parallelDG.graph.junction_tree as jtlib
    JM = jtlib.JunctionMap
    if not getattr(JM, "_pdg_shimmed", False):
        _orig_create = JM.create_t_and_t2clique
        _orig_connect = JM.connect
        _orig_disconnect = JM.disconnect

        def create_t_and_t2clique(self, jt, p=None):
            t, t2clique = _orig_create(self, jt, p)
            t2 = nx.Graph()
            t2.add_nodes_from(int(x) for x in t.nodes())
            t2.add_edges_from((int(a), int(b)) for a, b in t.edges())
            return t2, {int(k): v for k, v in t2clique.items()}

        def connect(self, t_node, g_node):
            return _orig_connect(self, int(t_node), g_node)

        def disconnect(self, t_node, g_node):
            return _orig_disconnect(self, int(t_node), g_node)

        JM.create_t_and_t2clique = create_t_and_t2clique
        JM.connect = connect
        JM.disconnect = disconnect
        JM._pdg_shimmed = True

    import parallelDG.mh_parallel as mh
    import parallelDG.graph.parallel_moves as ndlib
    import parallelDG.graph.decomposable as dlib
    import parallelDG.distributions.sequential_junction_tree_distributions as seqdist
    import parallelDG.distributions.discrete_dec_log_linear as loglin
    import parallelDG.distributions.gaussian_graphical_model as ggm
    import parallelDG.distributions.g_intra_class as gic
    ns = types.SimpleNamespace(mh=mh, ndlib=ndlib, jtlib=jtlib, dlib=dlib, seqdist=seqdist,
                               loglin=loglin, ggm=ggm, wish=wish, gic=gic, nx=nx)
    _loaded = ns
    return ns


def bits_of(sets, p, W):
    """list of p python sets -> (p, W) uint64 bit matrix."""
    out = np.zeros((p, W), dtype=np.uint64)
    for t in range(p):
        for g in sets[t]:
            out[t, g >> 6] |= np.uint64(1) << np.uint64(g & 63)
    return out


def set_to_bits(s, W):
    out = np.zeros((W,), dtype=np.uint64)
    for g in s:
        out[int(g) >> 6] |= np.uint64(1) << np.uint64(int(g) & 63)
    return out


class _Tracer(object):
    """Collects the replay log of one chain by wrapping the RNG, the proposal functions, the
    plugin objects and JunctionMap (tracing does not perturb the chain)."""

    def __init__(self, ref, p, single):
        self.ref = ref
        self.p = p
        self.W = (p + 63) // 64
        self.single = single
        self.states = []       # (iteration, edges, bits)
        self.it = []           # per iteration dict
        self.cur = None
        self.pending_lik = []
        self.pending_prior = []
        self.pending_logarg = []
        self.iter_index = 0

    def snap_state(self, jm, iteration):
        edges = np.array(sorted((min(a, b), max(a, b)) for a, b in jm.t.edges()), dtype=np.int32)
        edges = edges.reshape(-1, 2)
        bits = bits_of([jm.t2clique[k] for k in range(self.p)], self.p, self.W)
        self.states.append((iteration, edges, bits))


def trace_chain(kind, n_samples, randomize, sd, sd_graph, seed, init_graph=None):
    """Runs ONE reference chain (`kind` = 'parallel' -> mh_parallel.py:182-315 or
    'single' -> mh_parallel.py:23-180) and returns the replay log as a dict of numpy arrays."""
    ref = load_reference()
    mh, ndlib, jtlib = ref.mh, ref.ndlib, ref.jtlib
    p = sd.p
    single = (kind == "single")
    tr = _Tracer(ref, p, single)
    W = tr.W

    class SD(object):
        def __init__(self, inner):
            self.__dict__["_inner"] = inner

        def __getattr__(self, k):
            return getattr(self._inner, k)

        def __setattr__(self, k, v):
            setattr(self._inner, k, v)

        def log_likelihood_partial(self, cliques, separators):
            v = self._inner.log_likelihood_partial(cliques, separators)
            tr.pending_lik.append(float(v))
            return v

        def __str__(self):
            return str(self._inner)

    class SDG(object):
        def __init__(self, inner):
            self.__dict__["_inner"] = inner

        def __getattr__(self, k):
            return getattr(self._inner, k)

        def log_prior_partial(self, clq, sep, extra_arg=None):
            v = self._inner.log_prior_partial(clq, sep, extra_arg)
            tr.pending_prior.append(float(v))
            return v

    saved = dict(randint=np.random.randint, uniform=np.random.uniform,
                 conn=ndlib.propose_connect_moves_set, disc=ndlib.propose_disconnect_moves_set,
                 jm_init=jtlib.JunctionMap.__init__, jm_rand=jtlib.JunctionMap.randomize_by_jt,
                 shuffle=np.random.shuffle, tqdm=mh.tqdm, mhnp=mh.np)
    state = {"phase": 0, "jm": None, "in_loop": False}

    def randint(*a, **k):
        v = saved["randint"](*a, **k)
        if state["in_loop"] and not state.get("in_rand", False):
            if state["phase"] == 0:
                tr.cur = dict(node=int(v), mtype_raw=-1, moves=[], aux=-1, kind=-1, m=-1,
                              chosen=None, num_moves=0)
                tr.it.append(tr.cur)
                state["phase"] = 1
            elif state["phase"] == 1:
                tr.cur["mtype_raw"] = int(v)
                state["phase"] = 2
        return v

    def conn(tree, subtree):
        moves, num = saved["conn"](tree, subtree)
        if state["in_loop"] and state["phase"] == 2:
            # forward proposal of a connect iteration
            tr.cur["kind"] = 0
            tr.cur["m"] = len(subtree)
            tr.cur["moves"] = [(int(U), int(Ua) if isinstance(Ua, int) else -1) for U, Ua in moves]
            tr.cur["num_moves"] = num
            if len(subtree) == 0:
                tr.cur["aux"] = int(moves[0][0])
            state["phase"] = 3
            tr.pending_lik[:] = []
            tr.pending_prior[:] = []
            tr.pending_logarg[:] = []
        return moves, num

    def disc(tree, subtree, all_leafs=False):
        moves, num = saved["disc"](tree, subtree, all_leafs)
        if state["in_loop"] and state["phase"] == 2 and not all_leafs:
            tr.cur["kind"] = 1
            tr.cur["m"] = len(subtree)
            tr.cur["moves"] = [(int(U), int(Ua) if isinstance(Ua, int) else -1) for U, Ua in moves]
            tr.cur["num_moves"] = num
            if len(subtree) == 2:
                tr.cur["aux"] = int(moves[0][0])
            state["phase"] = 3
            tr.pending_lik[:] = []
            tr.pending_prior[:] = []
            tr.pending_logarg[:] = []
        return moves, num

    def shuffle(x):
        saved["shuffle"](x)
        if state["in_loop"] and state["phase"] == 3 and single and not state.get("in_rand", False):
            # mh_parallel.py:87/128 -- only moves[0] is used afterwards
            if len(x) and isinstance(x[0], tuple):
                U, Ua = x[0]
                tr.cur["chosen"] = (int(U), int(Ua) if isinstance(Ua, int) else -1)

    def uniform(*a, **k):
        v = saved["uniform"](*a, **k)
        if state["in_loop"] and state["phase"] == 3:
            lik = list(tr.pending_lik)
            pri = list(tr.pending_prior)
            tr.pending_lik[:] = []
            tr.pending_prior[:] = []
            kind = tr.cur["kind"]
            # order of calls: log_p2, log_p1, [extra]; log_g2, log_g1, [extra]
            lp2, lp1 = lik[0], lik[1]
            lg2, lg1 = pri[0], pri[1]
            if len(lik) == 3:
                if kind == 0:
                    lp1 += lik[2]
                    lg1 += pri[2]
                else:
                    lp2 += lik[2]
                    lg2 += pri[2]
            la = list(tr.pending_logarg)
            tr.pending_logarg[:] = []
            if single:
                # log_q2 - log_q1 = -log(num_moves) + log(num_reverse)  (mh_parallel.py:111-118)
                logq = -float(np.log(la[0])) + float(np.log(la[1]))
                tr.cur.setdefault("nfwd", []).append(la[0])
                tr.cur.setdefault("nrev", []).append(la[1])
            else:
                logq = -float(np.log(2)) if (kind == 1 and tr.cur["m"] == 2) else 0.0
                tr.cur.setdefault("nfwd", []).append(0.0)
                tr.cur.setdefault("nrev", []).append(0.0)
            tr.cur.setdefault("logq", []).append(logq)
            tr.cur.setdefault("u", []).append(float(v))
            tr.cur.setdefault("dlik", []).append(lp2 - lp1)
            tr.cur.setdefault("dprior", []).append(lg2 - lg1)
            tr.cur.setdefault("terms", []).append(abs(lp2) + abs(lp1))
        return v

    def jm_init(self, junction_tree):
        saved["jm_init"](self, junction_tree)
        state["jm"] = self
        tr.snap_state(self, 0)

    def jm_rand(self):
        state["in_rand"] = True
        try:
            saved["jm_rand"](self)
        finally:
            state["in_rand"] = False
        tr.snap_state(self, state["iter"])

    def tqdm(rng, **k):
        state["in_loop"] = True
        for i in rng:
            state["iter"] = i
            state["phase"] = 0
            yield i
        state["in_loop"] = False

    class NPProxy(object):
        """mh_parallel's `np`: identical to numpy except that np.log records its argument
        (mh_parallel.py:111-112,154-155,277 are the only np.log call sites in the loops)."""
        def __getattr__(self, k):
            return getattr(saved["mhnp"], k)

        def log(self, x):
            if state["in_loop"] and state["phase"] == 3:
                tr.pending_logarg.append(float(x))
            return saved["mhnp"].log(x)

    mh.np = NPProxy()
    np.random.randint = randint
    np.random.uniform = uniform
    np.random.shuffle = shuffle
    ndlib.propose_connect_moves_set = conn
    ndlib.propose_disconnect_moves_set = disc
    jtlib.JunctionMap.__init__ = jm_init
    jtlib.JunctionMap.randomize_by_jt = jm_rand
    mh.tqdm = tqdm
    try:
        random.seed(seed)
        fn = mh.sample_trajectory_single_move if single else mh.sample_trajectory
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            traj = fn(n_samples, randomize, SD(sd), SDG(sd_graph), init_graph=init_graph, seed=seed)
    finally:
        np.random.randint = saved["randint"]
        np.random.uniform = saved["uniform"]
        np.random.shuffle = saved["shuffle"]
        ndlib.propose_connect_moves_set = saved["conn"]
        ndlib.propose_disconnect_moves_set = saved["disc"]
        jtlib.JunctionMap.__init__ = saved["jm_init"]
        jtlib.JunctionMap.randomize_by_jt = saved["jm_rand"]
        mh.tqdm = saved["tqdm"]
        mh.np = saved["mhnp"]

    # ---- flatten -------------------------------------------------------------------------
    M = n_samples
    it_node = np.full((M,), -1, np.int32)
    it_mtype_raw = np.full((M,), -1, np.int32)
    it_kind = np.full((M,), -1, np.int32)
    it_aux = np.full((M,), -1, np.int32)
    it_off = np.zeros((M + 1,), np.int64)
    mv_U, mv_Uadj, mv_u, mv_dlik, mv_dprior, mv_terms = [], [], [], [], [], []
    mv_logq, mv_nfwd, mv_nrev = [], [], []
    assert len(tr.it) == M - 1, (len(tr.it), M)
    for i, rec in enumerate(tr.it, start=1):
        it_node[i] = rec["node"]
        it_mtype_raw[i] = rec["mtype_raw"]
        it_kind[i] = rec["kind"]
        it_aux[i] = rec["aux"]
        moves = rec["moves"]
        if single:
            moves = [rec["chosen"]] if rec["chosen"] is not None else []
        us = rec.get("u", [])
        assert len(us) == len(moves), (i, len(us), len(moves))
        for j, (U, Ua) in enumerate(moves):
            mv_U.append(U)
            mv_Uadj.append(Ua)
            mv_u.append(us[j])
            mv_dlik.append(rec["dlik"][j])
            mv_dprior.append(rec["dprior"][j])
            mv_terms.append(rec["terms"][j])
            mv_logq.append(rec["logq"][j])
            mv_nfwd.append(rec["nfwd"][j])
            mv_nrev.append(rec["nrev"][j])
        it_off[i + 1] = len(mv_U)
    it_off[1] = 0
    ns = len(tr.states)
    state_iter = np.array([s[0] for s in tr.states], np.int32)
    state_edges = np.stack([s[1] for s in tr.states]).astype(np.int32)
    state_bits = np.stack([s[2] for s in tr.states])
    upd = traj.jt_updates
    upd_i = np.array([u[0] for u in upd], np.int32)
    upd_k = np.array([u[1] for u in upd], np.int32)
    upd_type = np.array([u[2] for u in upd], np.int32)
    upd_node = np.array([list(u[3])[0] for u in upd], np.int32)
    nu = len(upd)
    upd_Cnew = np.zeros((nu, W), np.uint64)
    upd_C = np.zeros((nu, W), np.uint64)
    upd_Cadj = np.zeros((nu, W), np.uint64)
    for r, u in enumerate(upd):
        upd_Cnew[r] = set_to_bits(u[4][0], W)
        upd_C[r] = set_to_bits(u[4][1], W)
        upd_Cadj[r] = set_to_bits(u[4][2], W)
    log = dict(
        p=np.int32(p), n_samples=np.int32(M), randomize=np.int32(randomize), seed=np.int64(seed),
        single=np.int32(1 if single else 0),
        state_iter=state_iter, state_edges=state_edges, state_bits=state_bits,
        it_node=it_node, it_mtype_raw=it_mtype_raw, it_kind=it_kind, it_aux=it_aux, it_off=it_off,
        mv_U=np.array(mv_U, np.int32), mv_Uadj=np.array(mv_Uadj, np.int32),
        mv_u=np.array(mv_u, np.float64), mv_dlik=np.array(mv_dlik, np.float64),
        mv_dprior=np.array(mv_dprior, np.float64), mv_terms=np.array(mv_terms, np.float64),
        mv_logq=np.array(mv_logq, np.float64), mv_nfwd=np.array(mv_nfwd, np.int32),
        mv_nrev=np.array(mv_nrev, np.int32),
        logl=np.array(traj.logl, np.float64), n_updates=np.int64(traj.n_updates),
        upd_i=upd_i, upd_k=upd_k, upd_type=upd_type, upd_node=upd_node,
        upd_Cnew=upd_Cnew, upd_C=upd_C, upd_Cadj=upd_Cadj,
    )
    return log, traj


def postprocess_shims():
    """Shims for the reference's OUTPUT code (trajectory.py:126-370, empirical_graph_distribution.py,
    graph/graph.py:86-95), which uses pandas / networkx / numpy calls removed upstream.
    Each restores the removed call with its documented behaviour; none touches the algorithm."""
    import networkx as nx
    import pandas as pd
    if not hasattr(nx, "to_numpy_matrix"):
        nx.to_numpy_matrix = lambda g, *a, **k: np.matrix(nx.to_numpy_array(g, *a, **k))
    if not hasattr(pd.DataFrame, "append"):
        pd.DataFrame.append = lambda self, other, **k: pd.concat([self, other], **k)
    if not getattr(pd.Series.fillna, "_pdg_shim", False):
        _orig = pd.Series.fillna

        def fillna(self, value=None, method=None, **k):
            if method == "ffill":
                return self.ffill()
            if method == "bfill":
                return self.bfill()
            return _orig(self, value, **k)
        fillna._pdg_shim = True
        pd.Series.fillna = fillna
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "float_"):
        np.float_ = np.float64


def run_chain(kind, n_samples, randomize, sd, sd_graph, seed, init_graph=None):
    """one UNTRACED reference chain (same seeding discipline as trace_chain); returns its Trajectory"""
    import contextlib
    import io
    ref = load_reference()
    random.seed(seed)
    fn = ref.mh.sample_trajectory_single_move if kind == "single" else ref.mh.sample_trajectory
    saved = ref.mh.tqdm
    ref.mh.tqdm = lambda rng, **k: rng
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            return fn(n_samples, randomize, sd, sd_graph, init_graph=init_graph, seed=seed)
    finally:
        ref.mh.tqdm = saved


def trace_greenthomas(n_samples, randomize, sd, seed):
    """One chain of the reference's Green-Thomas sampler (mh_greenthomas.py:20-233) with every proposal
    logged: the junction tree before the move (cliques, edges), the move's own choices (clique / edge,
    X, Y, the coin flips of disconnect case a), the case letter, both proposal log-probabilities, the
    log targets and the accept draw; the tree after every randomisation.  Logging wraps
    aglib.connect_move / disconnect_move / connect_logprob / disconnect_logprob_* and
    np.random.choice; it draws nothing itself, so the chain is the untraced one."""
    import contextlib
    import io
    ref = load_reference()
    postprocess_shims()
    import parallelDG.mh_greenthomas as gt
    import parallelDG.graph.greenthomas as aglib
    import parallelDG.graph.junction_tree as jtlib
    log = {"moves": {}, "randomized": {}}
    cur = {}

    def snap(tree):
        return [sorted(c) for c in tree.nodes()], [(sorted(a), sorted(b)) for a, b in tree.edges()]

    orig = dict(conn=aglib.connect_move, disc=aglib.disconnect_move, sel_c=aglib.connect_select_subsets,
                sel_d=aglib.disconnect_select_subsets, rs=aglib.aux.random_subset, clp=aglib.connect_logprob,
                dla=aglib.disconnect_logprob_a, dlb=aglib.disconnect_logprob_bcd, rand=jtlib.randomize,
                choice=np.random.choice, tqdm=gt.tqdm)
    state = {"i": 0, "in_move": False}

    def begin(kind, tree):
        nodes, edges = snap(tree)
        cur.clear()
        cur.update(kind=kind, nodes=nodes, edges=edges, case=None, X=[], Y=[], with_x=[], log_q12=None, log_q21=None)

    def connect_move(tree):
        begin("connect", tree)
        r = orig["conn"](tree)
        if r:
            cur["case"], cur["log_q12"] = r[0], float(r[1])
        return r

    def disconnect_move(tree):
        begin("disconnect", tree)
        r = orig["disc"](tree)
        if r is not False:
            cur["case"], cur["log_q12"] = r[0], float(r[1])
        return r

    def sel_c(tree):
        X, Y, CX, CY = orig["sel_c"](tree)
        cur.update(X=sorted(X), Y=sorted(Y), CX=sorted(CX), CY=sorted(CY))
        return X, Y, CX, CY

    def sel_d(tree, c):
        X, Y = orig["sel_d"](tree, c)
        cur.update(X=sorted(X), Y=sorted(Y), C=sorted(c))
        return X, Y

    def random_subset(A):
        r = orig["rs"](A)
        cur["with_x"] = [sorted(n) for n in r]
        return r

    # the reverse-move probability is whichever of these is called from sample_trajectory after the move
    def wrap_rev(f):
        def g(*a, **k):
            v = f(*a, **k)
            if state["in_move"] is False:
                cur["log_q21"] = float(v)
            return v
        return g

    def conn_guard(f):
        def g(tree):
            state["in_move"] = True
            try:
                return f(tree)
            finally:
                state["in_move"] = False
        return g

    def randomize_tree(tree):
        orig["rand"](tree)
        log["randomized"][state["i"]] = snap(tree)

    def choice(a, size=None, replace=True, p=None):
        r = orig["choice"](a, size=size, replace=replace, p=p)
        if p is not None and isinstance(a, (int, np.integer)) and a == 2 and len(p) == 2:
            cur["accepted"] = int(np.asarray(r).ravel()[0])
            cur["alpha"] = float(p[1])
        return r

    class Seq(object):
        """the iteration range: tells the tracer the MCMC index and files the finished proposal"""
        def __init__(self, rng):
            self.rng = rng

        def __iter__(self):
            for i in self.rng:
                state["i"] = i
                cur.clear()
                cur.update(kind=None, case=None)
                yield i
                log["moves"][i] = dict(cur)

    aglib.connect_move = conn_guard(connect_move)
    aglib.disconnect_move = conn_guard(disconnect_move)
    aglib.connect_select_subsets = sel_c
    aglib.disconnect_select_subsets = sel_d
    aglib.aux.random_subset = random_subset
    aglib.connect_logprob = wrap_rev(orig["clp"])
    aglib.disconnect_logprob_a = wrap_rev(orig["dla"])
    aglib.disconnect_logprob_bcd = wrap_rev(orig["dlb"])
    jtlib.randomize = randomize_tree
    np.random.choice = choice
    gt.tqdm = lambda rng, **k: Seq(rng)
    # log target after the move = sd.log_likelihood(jt) - log_n_junction_trees(jt, seps)  (:67, :73, :146)
    orig_ll, orig_mu = sd.log_likelihood, jtlib.log_n_junction_trees

    def log_likelihood(jt):
        v = orig_ll(jt)
        cur["ll_after"] = float(v)
        return v

    def log_mu(tree, S):
        v = orig_mu(tree, S)
        cur["mu_after"] = float(v)
        return v
    sd.log_likelihood = log_likelihood
    jtlib.log_n_junction_trees = log_mu
    try:
        np.random.seed(seed)
        random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            traj = gt.sample_trajectory(n_samples, randomize, sd, seed=seed)
    finally:
        aglib.connect_move, aglib.disconnect_move = orig["conn"], orig["disc"]
        aglib.connect_select_subsets, aglib.disconnect_select_subsets = orig["sel_c"], orig["sel_d"]
        aglib.aux.random_subset = orig["rs"]
        aglib.connect_logprob, aglib.disconnect_logprob_a, aglib.disconnect_logprob_bcd = orig["clp"], orig["dla"], orig["dlb"]
        jtlib.randomize = orig["rand"]
        np.random.choice = orig["choice"]
        gt.tqdm = orig["tqdm"]
        del sd.log_likelihood
        jtlib.log_n_junction_trees = orig_mu
    log["logl"] = [float(x) for x in traj.logl]
    log["graphs"] = [sorted(tuple(sorted((int(a), int(b)))) for a, b in g.edges()) for g in traj.trajectory]
    for i, mv in log["moves"].items():
        mv["log_p1"] = log["logl"][i - 1]
        if "ll_after" in mv:
            mv["log_p2"] = mv["ll_after"] - mv["mu_after"]
    return log, traj
