"""TEST INFRASTRUCTURE ONLY -- ctypes front-end for the two CPU oracles.

* ``load_port()``  -> oracle/libbb_oracle.so   (plain-C restatement, oracle/bb_oracle.c, prefix ``orc_``)
* ``load_ref()``   -> oracle/_ref/libdgref.so  (UNMODIFIED reference sources + oracle/ref_shim.cpp, prefix ``ref_``)

Both expose the same flat API, so every test can be run against either.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module;
nothing under ``deepgroebner_b200/`` does.

Python-level polynomial format used throughout the tests ("tuple polys"):
    poly  = [(coef, (e0, e1, ..., e7)), ...]      # descending grevlex, coef in [0, P)
    ideal = [poly, ...]
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NV = 8
W = 9

ELIM = {"gebauermoeller": 0, "lcm": 1, "none": 2}
REWARDS = {"additions": 0, "reductions": 1}
SELECT = {"first": 0, "degree": 1, "normal": 2, "sugar": 3, "random": 4, "last": 5, "codegree": 6, "strange": 7,
          "spice": 8}

_i = C.c_int
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_cp = C.c_char_p


def _np_i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    return a.ctypes.data_as(_ip)


def pad(e):
    e = tuple(int(x) for x in e)
    return e + (0,) * (NV - len(e))


def to_wire(poly):
    """tuple poly -> int32 [n*9]"""
    out = np.zeros((len(poly), W), dtype=np.int32)
    for r, (c, e) in enumerate(poly):
        out[r, 0] = c
        out[r, 1:] = pad(e)
    return out.reshape(-1)


def from_wire(buf, n):
    a = np.asarray(buf[: n * W]).reshape(n, W)
    return [(int(r[0]), tuple(int(x) for x in r[1:])) for r in a]


def ideal_to_wire(F):
    lens = _np_i([len(f) for f in F])
    terms = np.concatenate([to_wire(f) for f in F]) if len(F) and sum(len(f) for f in F) else np.zeros(0, np.int32)
    return _np_i(terms), lens


def ideal_from_wire(terms, lens, npoly):
    out, off = [], 0
    for p in range(npoly):
        n = int(lens[p])
        out.append(from_wire(terms[off * W:], n))
        off += n
    return out


class Oracle:
    """Thin, allocation-happy wrapper: correctness tooling, not a fast path."""

    CAP_TERMS = 1 << 18
    CAP_POLYS = 1 << 14
    CAP_PAIRS = 1 << 16

    def __init__(self, path, prefix):
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path)
        self.kind = "reference" if prefix == "ref_" else "port"
        L = self._f
        L("prime", _i)
        L("nslots", _i)
        for n in ("coef_div", "coef_mul", "coef_add", "coef_sub"):
            L(n, _i, _i, _i)
        L("coef_norm", _i, _i)
        L("mono_cmp", _i, _ip, _ip)
        L("mono_divisible", _i, _ip, _ip)
        L("mono_lcm", None, _ip, _ip, _ip)
        L("poly_make", _i, _ip, _i, _ip, _i)
        for n in ("poly_add", "poly_sub", "poly_mul", "spoly"):
            L(n, _i, _ip, _i, _ip, _i, _ip, _i)
        L("term_mul", _i, _ip, _ip, _i, _ip, _i)
        L("parse_polynomial", _i, _cp, _ip, _i)
        L("reduce", _i, _ip, _i, _ip, _ip, _i, _ip, _i, _ip)
        L("update", _i, _ip, _ip, _i, _ip, _i, _i, _ip, _i, _i)
        L("minimalize", _i, _ip, _ip, _i, _ip, _i, _ip, _i)
        L("interreduce", _i, _ip, _ip, _i, _ip, _i, _ip, _i)
        L("buchberger", _i, _ip, _ip, _i, _i, _i, _i, _i, _i, C.c_double, _i, _ip, _i, _ip, _i, _dp)
        L("gen_create", _vp, _cp)
        L("gen_destroy", None, _vp)
        L("gen_seed", None, _vp, _i)
        L("gen_nvars", _i, _vp)
        L("gen_next", _i, _vp, _ip, _i, _ip, _i)
        L("basis", _i, _i, _i, _ip, _i)
        L("degree_distribution", _i, _i, _i, _i, _i, _dp, _i)
        L("cyclic", _i, _i, _ip, _i, _ip, _i)
        L("env_create", _vp, _cp, _i, _i, _i, _i)
        L("env_destroy", None, _vp)
        L("env_seed", None, _vp, _i)
        L("env_set_ideal", None, _vp, _ip, _ip, _i)
        L("env_nvars", _i, _vp)
        L("env_reset", None, _vp)
        L("env_step", C.c_double, _vp, _i, _i)
        L("env_npairs", _i, _vp)
        L("env_nbasis", _i, _vp)
        L("env_nterms", _i, _vp)
        L("env_pairs", _i, _vp, _ip, _i)
        L("env_basis", _i, _vp, _ip, _i, _ip, _i)
        L("env_reducers", _i, _vp, _ip, _i, _ip, _i)
        L("env_value", C.c_double, _vp, _cp, C.c_double)
        L("env_value_seeded", C.c_double, _vp, _i, C.c_double, _i, _i)
        L("env_select", _i, _vp, _i)
        L("env_final_gb", _i, _vp, _ip, _i, _ip, _i)
        L("env_run", _i, _vp, _i, _ip, _i, _ip, _i)
        L("lm_create", _vp, _cp, _i, _i, _i)
        L("lm_destroy", None, _vp)
        L("lm_seed", None, _vp, _i)
        L("lm_set_ideal", None, _vp, _ip, _ip, _i, _i)
        L("lm_reset", None, _vp)
        L("lm_step", C.c_double, _vp, _i)
        L("lm_cols", _i, _vp)
        L("lm_state", _i, _vp, _ip, _i)
        L("lm_value", C.c_double, _vp, _cp, C.c_double)
        L("bench_selection", None, _cp, _i, _i, _i, _i, _i, _i, _dp)
        if prefix == "ref_":
            L("bench_random", None, _cp, _i, _i, _i, _dp)
            L("bench_buchberger", None, _cp, _i, _i, _i, _i, _i, C.c_double, _dp)
            L("run_records", _i, _cp, _i, _i, _i, _i, _i, _i, _ip, _i, _i, _i, _i, C.c_double, _i, _i, _vp)
        else:
            L("set_prime", None, _i)

    def _f(self, name, res, *args):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = res
        fn.argtypes = list(args)
        setattr(self, "c_" + name, fn)

    # ---- scalars / monomials
    def prime(self):
        return self.c_prime()

    def set_prime(self, p):
        if self.prefix == "ref_":
            if p != 32003:
                raise ValueError("the reference is fixed at P=32003")
        else:
            self.c_set_prime(p)

    def mono_cmp(self, a, b):
        a, b = _np_i(pad(a)), _np_i(pad(b))
        return self.c_mono_cmp(_ptr(a), _ptr(b))

    def mono_divisible(self, a, b):
        a, b = _np_i(pad(a)), _np_i(pad(b))
        return bool(self.c_mono_divisible(_ptr(a), _ptr(b)))

    def mono_lcm(self, a, b):
        a, b, o = _np_i(pad(a)), _np_i(pad(b)), np.zeros(NV, np.int32)
        self.c_mono_lcm(_ptr(a), _ptr(b), _ptr(o))
        return tuple(int(x) for x in o)

    # ---- polynomials
    def _poly_out(self, fn, *args, cap=4096):
        out = np.zeros(cap * W, np.int32)
        n = fn(*args, _ptr(out), cap)
        if n < 0:
            return self._poly_out(fn, *args, cap=-n + 8)
        return from_wire(out, n)

    def poly_make(self, f):
        a = to_wire(f)
        return self._poly_out(self.c_poly_make, _ptr(a), len(f))

    def _binop(self, fn, f, g):
        a, b = to_wire(f), to_wire(g)
        return self._poly_out(fn, _ptr(a), len(f), _ptr(b), len(g))

    def poly_add(self, f, g):
        return self._binop(self.c_poly_add, f, g)

    def poly_sub(self, f, g):
        return self._binop(self.c_poly_sub, f, g)

    def poly_mul(self, f, g):
        return self._binop(self.c_poly_mul, f, g)

    def spoly(self, f, g):
        return self._binop(self.c_spoly, f, g)

    def term_mul(self, t, f):
        tt, a = to_wire([t]), to_wire(f)
        return self._poly_out(self.c_term_mul, _ptr(tt), _ptr(a), len(f))

    def parse_polynomial(self, s):
        return self._poly_out(self.c_parse_polynomial, s.encode())

    def reduce(self, g, F):
        a = to_wire(g)
        ft, fl = ideal_to_wire(F)
        out = np.zeros(self.CAP_TERMS * W, np.int32)
        steps = C.c_int(0)
        n = self.c_reduce(_ptr(a), len(g), _ptr(ft), _ptr(fl), len(F), _ptr(out), self.CAP_TERMS, C.byref(steps))
        assert n >= 0
        return from_wire(out, n), steps.value

    def update(self, G, P, f, elimination="gebauermoeller"):
        gt, gl = ideal_to_wire(G)
        pairs = np.zeros(2 * self.CAP_PAIRS, np.int32)
        for r, (i, j) in enumerate(P):
            pairs[2 * r], pairs[2 * r + 1] = i, j
        a = to_wire(f)
        n = self.c_update(_ptr(gt), _ptr(gl), len(G), _ptr(pairs), len(P), self.CAP_PAIRS, _ptr(a), len(f),
                          ELIM[elimination])
        assert n >= 0
        return [(int(pairs[2 * r]), int(pairs[2 * r + 1])) for r in range(n)]

    def _ideal_out(self, fn, *args):
        ot = np.zeros(self.CAP_TERMS * W, np.int32)
        ol = np.zeros(self.CAP_POLYS, np.int32)
        n = fn(*args, _ptr(ot), self.CAP_TERMS, _ptr(ol), self.CAP_POLYS)
        if n < 0:
            raise RuntimeError("oracle call failed (%d)" % n)
        return ideal_from_wire(ot, ol, n)

    def minimalize(self, G):
        gt, gl = ideal_to_wire(G)
        return self._ideal_out(self.c_minimalize, _ptr(gt), _ptr(gl), len(G))

    def interreduce(self, G):
        gt, gl = ideal_to_wire(G)
        return self._ideal_out(self.c_interreduce, _ptr(gt), _ptr(gl), len(G))

    def buchberger(self, F, selection="degree", elimination="gebauermoeller", rewards="additions", sort_input=False,
                   sort_reducers=True, gamma=0.99, seed=0):
        ft, fl = ideal_to_wire(F)
        ot = np.zeros(self.CAP_TERMS * W, np.int32)
        ol = np.zeros(self.CAP_POLYS, np.int32)
        st = np.zeros(5, np.float64)
        n = self.c_buchberger(_ptr(ft), _ptr(fl), len(F), SELECT[selection], ELIM[elimination], REWARDS[rewards],
                              int(sort_input), int(sort_reducers), gamma, seed, _ptr(ot), self.CAP_TERMS, _ptr(ol),
                              self.CAP_POLYS, st.ctypes.data_as(_dp))
        assert n >= 0
        stats = dict(zero_reductions=int(st[0]), nonzero_reductions=int(st[1]), polynomial_additions=int(st[2]),
                     total_reward=float(st[3]), discounted_return=float(st[4]))
        return ideal_from_wire(ot, ol, n), stats

    # ---- generators
    def basis(self, n, d):
        out = np.zeros(NV * 100000, np.int32)
        m = self.c_basis(n, d, _ptr(out), 100000)
        assert m >= 0
        return [tuple(int(x) for x in out[i * NV:(i + 1) * NV]) for i in range(m)]

    def degree_distribution(self, n, d, dist="uniform", constants=False):
        out = np.zeros(d + 2, np.float64)
        m = self.c_degree_distribution(n, d, {"uniform": 0, "weighted": 1, "maximum": 2}[dist], int(constants),
                                       out.ctypes.data_as(_dp), d + 2)
        return [float(x) for x in out[:m]]

    def cyclic(self, n):
        return self._ideal_out(self.c_cyclic, n)

    def generator(self, dist):
        return Generator(self, dist)

    def env(self, dist="3-20-10-uniform", elimination="gebauermoeller", rewards="additions", sort_input=False,
            sort_reducers=True):
        return Env(self, dist, elimination, rewards, sort_input, sort_reducers)

    def lm_env(self, dist="3-20-10-uniform", sort_input=False, sort_reducers=True, k=2):
        return LmEnv(self, dist, sort_input, sort_reducers, k)

    def bench_selection(self, dist, selection, seed0, count, nthreads=1, with_matrix=False, k=2):
        out = np.zeros(3, np.float64)
        self.c_bench_selection(dist.encode(), SELECT[selection], seed0, count, nthreads, int(with_matrix), k,
                               out.ctypes.data_as(_dp))
        return dict(steps=int(out[0]), additions=int(out[1]), seconds=float(out[2]))

    def bench_buchberger(self, dist, selection, seed0, count, nthreads=1, sel_seed0=0, gamma=0.99):
        """Whole episodes through the reference's buchberger() loop (any SelectionType, seeded Random); ref only."""
        out = np.zeros(3, np.float64)
        self.c_bench_buchberger(dist.encode(), SELECT[selection], seed0, count, nthreads, sel_seed0, gamma,
                                out.ctypes.data_as(_dp))
        return dict(steps=int(out[0]), additions=int(out[1]), seconds=float(out[2]))

    # numpy view of one record = bb_episode_stats (include/bbenv.h); kept here so that oracle/ imports nothing of the product
    RECORD_DTYPE = [("steps", "<i4"), ("additions", "<i4"), ("zero_reductions", "<i4"), ("nonzero_reductions", "<i4"),
                    ("nbasis", "<i4"), ("nterms", "<i4"), ("status", "<i4"), ("rerolls", "<i4"), ("trace_hash", "<u8"),
                    ("basis_hash", "<u8"), ("gb_hash", "<u8"), ("gb_polys", "<i4"), ("gb_terms", "<i4"),
                    ("discounted_return", "<f8")]

    def run_records(self, dist, selection, count, seed0=0, seeds=None, sel_seed0=0, sel_stride=1, max_steps=0,
                    gamma=0.99, compute_gb=True, nthreads=0, elimination="gebauermoeller", sort_input=False,
                    sort_reducers=True):
        """Per-episode records (the layout bb_run writes) of episodes 0..count-1 from the unmodified reference env,
        on `nthreads` host threads (0 = all): the exhaustive parity gate of bench.py and the full-size GPU tests."""
        if self.prefix != "ref_":
            raise RuntimeError("run_records needs oracle/_ref (the unmodified reference)")
        if not nthreads:
            nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        out = np.zeros(count, dtype=np.dtype(self.RECORD_DTYPE))
        sd = None if seeds is None else _np_i(seeds)
        rc = self.c_run_records(dist.encode(), SELECT[selection], ELIM[elimination], 0, int(sort_input),
                                int(sort_reducers), seed0, None if sd is None else _ptr(sd), count, sel_seed0, sel_stride,
                                max_steps, gamma, int(compute_gb), nthreads, out.ctypes.data_as(_vp))
        if rc < 0:
            raise RuntimeError("ref_run_records failed (%d)" % rc)
        return out

    def bench_random(self, dist, seed0, episodes, nthreads=1):
        out = np.zeros(3, np.float64)
        self.c_bench_random(dist.encode(), seed0, episodes, nthreads, out.ctypes.data_as(_dp))
        return dict(steps=int(out[0]), additions=int(out[1]), seconds=float(out[2]))


class Generator:
    def __init__(self, orc, dist):
        self.o = orc
        self.h = orc.c_gen_create(dist.encode())
        if not self.h:
            raise ValueError("bad ideal_dist %r" % dist)

    def seed(self, s):
        self.o.c_gen_seed(self.h, s)

    def nvars(self):
        return self.o.c_gen_nvars(self.h)

    def next(self):
        return self.o._ideal_out(self.o.c_gen_next, self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.o.c_gen_destroy(self.h)
            self.h = None


class Env:
    """Single-episode BuchbergerEnv (buchberger.cpp:269-351 / buchberger.py:243-394) on the oracle."""

    def __init__(self, orc, dist, elimination, rewards, sort_input, sort_reducers):
        self.o = orc
        self.h = orc.c_env_create(dist.encode(), ELIM[elimination], REWARDS[rewards], int(sort_input),
                                  int(sort_reducers))
        if not self.h:
            raise ValueError("bad ideal_dist %r" % dist)

    def __del__(self):
        if getattr(self, "h", None):
            self.o.c_env_destroy(self.h)
            self.h = None

    def seed(self, s):
        self.o.c_env_seed(self.h, s)

    def set_ideal(self, F):
        ft, fl = ideal_to_wire(F)
        self.o.c_env_set_ideal(self.h, _ptr(ft), _ptr(fl), len(F))

    def nvars(self):
        return self.o.c_env_nvars(self.h)

    def reset(self):
        self.o.c_env_reset(self.h)
        return self.basis(), self.pairs()

    def step(self, action):
        i, j = action
        reward = self.o.c_env_step(self.h, i, j)
        return reward, self.o.c_env_npairs(self.h) == 0

    def pairs(self):
        n = self.o.c_env_npairs(self.h)
        buf = np.zeros(2 * max(n, 1), np.int32)
        self.o.c_env_pairs(self.h, _ptr(buf), n)
        return [(int(buf[2 * r]), int(buf[2 * r + 1])) for r in range(n)]

    def basis(self):
        return self.o._ideal_out(self.o.c_env_basis, self.h)

    def reducers(self):
        return self.o._ideal_out(self.o.c_env_reducers, self.h)

    def final_gb(self):
        return self.o._ideal_out(self.o.c_env_final_gb, self.h)

    def value(self, strategy="degree", gamma=0.99):
        return self.o.c_env_value(self.h, strategy.encode(), gamma)

    def value_seeded(self, strategy="degree", gamma=0.99, seed=0, rollouts=None):
        """value() with explicit seeds for the random strategies; strategy may be 'sample'.  rollouts None = the
        reference's count: 1 Degree + 100 Random for 'sample' (buchberger.cpp:333-341), 1 otherwise."""
        code = 100 if strategy == "sample" else SELECT[strategy]
        if rollouts is None:
            rollouts = 0 if strategy == "sample" else 1
        return self.o.c_env_value_seeded(self.h, code, gamma, seed, rollouts)

    def select(self, selection):
        return self.o.c_env_select(self.h, SELECT[selection])

    def run(self, selection=None, actions=None, cap=1 << 16):
        """Run to completion from the current state; returns int32 trace [T,5] = (i, j, additions, |P|, |G|)."""
        trace = np.zeros(5 * cap, np.int32)
        if selection is not None:
            n = self.o.c_env_run(self.h, SELECT[selection], None, 0, _ptr(trace), cap)
        else:
            a = _np_i(actions)
            n = self.o.c_env_run(self.h, -1, _ptr(a), len(a), _ptr(trace), cap)
        if n < 0:
            raise RuntimeError("env_run failed (%d)" % n)
        return trace[: 5 * n].reshape(n, 5).copy()


class LmEnv:
    """LeadMonomialsEnv (buchberger.cpp:373-408) as wrapped.pyx drives it."""

    def __init__(self, orc, dist, sort_input, sort_reducers, k):
        self.o = orc
        self.h = orc.c_lm_create(dist.encode(), int(sort_input), int(sort_reducers), k)
        if not self.h:
            raise ValueError("bad ideal_dist %r" % dist)

    def __del__(self):
        if getattr(self, "h", None):
            self.o.c_lm_destroy(self.h)
            self.h = None

    def seed(self, s):
        self.o.c_lm_seed(self.h, s)

    def set_ideal(self, F, nvars=0):
        ft, fl = ideal_to_wire(F)
        self.o.c_lm_set_ideal(self.h, _ptr(ft), _ptr(fl), len(F), nvars)

    def state(self):
        cols = self.o.c_lm_cols(self.h)
        cap = 1 << 20
        buf = np.zeros(cap, np.int32)
        n = self.o.c_lm_state(self.h, _ptr(buf), cap)
        assert n >= 0
        return buf[:n].reshape(-1, cols).copy()

    def reset(self):
        self.o.c_lm_reset(self.h)
        return self.state()

    def step(self, action):
        r = self.o.c_lm_step(self.h, int(action))
        s = self.state()
        return s, r, s.shape[0] == 0, {}

    def value(self, strategy="degree", gamma=0.99):
        return self.o.c_lm_value(self.h, strategy.encode(), gamma)


PORT_SO = os.path.join(HERE, "libbb_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libdgref.so")


def build(ref=True):
    """Compile the checkers (building is not using).  The reference build is skipped when /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", HERE, "port"] + (["ref"] if ref else []), check=True)


REF_MAKE_STRAT = os.path.join(HERE, "_ref", "make_strat")   # the reference's scripts/make_strat.cpp, unmodified


def load_port():
    if not os.path.exists(PORT_SO):
        build(ref=False)
    return Oracle(PORT_SO, "orc_")


def have_ref():
    return os.path.exists(REF_SO)


def load_ref():
    if not have_ref():
        raise FileNotFoundError(REF_SO + " (run `make -C oracle ref` where /root/reference exists)")
    return Oracle(REF_SO, "ref_")
