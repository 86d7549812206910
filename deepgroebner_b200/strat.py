"""Strategy statistics over a file of ideals: the contract of the reference's scripts/make_strat.cpp, one launch.

The reference reads ``data/stats/<dist>/<dist>.csv`` (a header line, then one ideal per line as polynomials separated
by ``|``, e.g. ``a^2*b+c*d|b^3-7*a``; written by scripts/make_dist.m2:62-72), runs ``buchberger(F, strategy,
GebauerMoeller, Additions, false, true, 0.99, seed)`` on every line (make_strat.cpp:64-70) and writes
``data/stats/<dist>/<dist>_<strategy>[_<seed>].csv`` with the columns ``ZeroReductions,NonzeroReductions,
PolynomialAdditions``.  Here every ideal of the file is staged on the device (bb_set_ideals) and all of them run to
completion in one bb_run launch; the output file is byte-identical to the reference binary's.
"""
import os

STRATEGIES = ("first", "degree", "normal", "sugar", "random", "last", "codegree", "strange", "spice")
HEADER = "ZeroReductions,NonzeroReductions,PolynomialAdditions"
NVARS_MAX = 8  # polynomials.h:29


def parse_polynomial(s, prime=32003):
    """parse_polynomial (polynomials.cpp:226-300): variables a..h, ``^`` powers, ``*`` products, integer
    coefficients, ``+``/``-`` between terms, no spaces.  Returns [(coef in [1,p), exponent 8-tuple)] in the order
    written, equal monomials summed (Polynomial operator+, polynomials.cpp:148-177) and zero sums dropped."""
    i, n = 0, len(s)

    def monomial():
        nonlocal i
        e = [0] * NVARS_MAX
        while True:
            if i >= n:
                return e
            v = ord(s[i]) - ord("a")
            if v < 0 or v >= NVARS_MAX:
                raise ValueError("invalid variable name %r in %r" % (s[i], s))
            i += 1
            power = 1
            if i < n and s[i] == "^":
                i += 1
                j = i
                while j < n and s[j].isdigit():
                    j += 1
                if j == i:
                    raise ValueError("missing exponent in %r" % s)
                power, i = int(s[i:j]), j
            e[v] += power
            if i < n and s[i] == "*":
                i += 1
                continue
            return e

    terms = {}
    order = []
    while i < n:
        sign = 1
        while i < n and s[i] in "+-":
            if s[i] == "-":
                sign = -sign
            i += 1
        if i < n and s[i].isdigit():
            j = i
            while j < n and s[j].isdigit():
                j += 1
            c, i = int(s[i:j]), j
            if i < n and s[i] == "*":
                i += 1
                e = monomial()
            else:
                e = [0] * NVARS_MAX
        else:
            c, e = 1, monomial()
        e = tuple(e)
        if e not in terms:
            order.append(e)
            terms[e] = 0
        terms[e] = (terms[e] + sign * c) % prime
    return [(terms[e], e) for e in order if terms[e]]


def parse_ideal_string(line, prime=32003):
    """parse_ideal_string (make_strat.cpp:12-19): polynomials separated by '|'."""
    return [parse_polynomial(p, prime) for p in line.split("|")]


def read_ideal_file(path, prime=32003):
    with open(path) as fh:
        lines = fh.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return [parse_ideal_string(line, prime) for line in lines[1:]]  # the first line is the column name


def strategy_stats(ideals, strategy, seed=0, device="cuda:0", capacity=None, **caps):
    """(zero_reductions, nonzero_reductions, polynomial_additions) int arrays for a list of ideals: one bb_run launch
    with every ideal staged on the device.  `seed`: the Random-selection seed shared by every ideal
    (make_strat.cpp:66 passes the same optional seed to every buchberger() call)."""
    from .buchberger import BuchbergerEngine
    from .ideals import FixedIdealGenerator
    if strategy not in STRATEGIES:
        raise ValueError("unknown strategy %r" % strategy)
    n = 1
    for F in ideals:
        for f in F:
            if not f:
                raise ValueError("the zero polynomial is not a valid generator")
            for _, e in f:
                for v, x in enumerate(e):
                    if x:
                        n = max(n, v + 1)
    binomial = all(len(f) <= 2 for F in ideals for f in F)
    eng = BuchbergerEngine(FixedIdealGenerator(ideals[0], n), num_envs=len(ideals), device=device,
                           capacity=capacity or ("binomial" if binomial else "poly"),
                           max_gens=max(len(F) for F in ideals),
                           max_gen_terms=max(sum(len(f) for f in F) for F in ideals), **caps)
    eng.set_ideals(ideals)
    eng.set_selection_seed_stride(0)
    stats, _ = eng.run_episodes(strategy, episodes=len(ideals), selection_seed=int(seed or 0), gamma=0.99)
    bad = (stats["status"] != 2).nonzero()[0]
    if len(bad):
        raise RuntimeError("ideal %d did not finish (status %d): raise the arena capacities (capacity='general' or "
                           "max_basis/max_pairs/max_terms keywords)" % (int(bad[0]), int(stats["status"][bad[0]])))
    return stats["zero_reductions"], stats["nonzero_reductions"], stats["additions"]


def make_strat(dist, strategy, seed=None, root="data/stats", device="cuda:0", **caps):
    """scripts/make_strat.cpp main(): same input/output file names, same refusal to overwrite, same CSV bytes.
    Returns (exit code, message): 0 ok, 2 no distribution file, 3 output exists (make_strat.cpp:35-48)."""
    in_name = os.path.join(root, dist, dist + ".csv")
    if not os.path.exists(in_name):
        return 2, "No distribution file found. Run scripts/make_dist.m2 first."
    out_name = os.path.join(root, dist, "%s_%s.csv" % (dist, strategy))
    if seed is not None and strategy == "random":
        out_name = os.path.join(root, dist, "%s_%s_%s.csv" % (dist, strategy, seed))
    if os.path.exists(out_name):
        return 3, "Output file %s already exists. Delete or move it first." % out_name
    if strategy == "random" and seed is None:
        raise ValueError("random selection needs a seed here (the reference falls back to std::random_device)")
    ideals = read_ideal_file(in_name)
    rows = []
    if ideals:
        z, nz, adds = strategy_stats(ideals, strategy, seed=seed, device=device, **caps)
        rows = ["%d,%d,%d" % (int(a), int(b), int(c)) for a, b, c in zip(z, nz, adds)]
    with open(out_name, "w") as fh:
        fh.write("\n".join([HEADER] + rows) + "\n")
    return 0, out_name


def format_polynomial(f, prime=32003):
    """A polynomial [(coef, exps), ...] in the file's notation (what Macaulay2's toString prints and
    parse_polynomial reads back): ``413*a^2*b^5*c+32*d^2-5``; coefficients above p/2 are written negative."""
    out = []
    for c, e in f:
        c = int(c) % prime
        if c > prime // 2:
            c -= prime
        mono = "*".join(("%s^%d" % (chr(97 + v), x) if x > 1 else chr(97 + v)) for v, x in enumerate(e) if x)
        if not mono:
            body = str(abs(c))
        elif abs(c) == 1:
            body = mono
        else:
            body = "%d*%s" % (abs(c), mono)
        out.append(("-" if c < 0 else ("+" if out else "")) + body)
    return "".join(out) if out else "0"


def write_ideal_file(path, ideals, prime=32003):
    """data/stats/<dist>/<dist>.csv as scripts/make_dist.m2:52-72 writes it: 'Ideal', then one ideal per line."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        fh.write("Ideal\n")
        for F in ideals:
            fh.write("|".join(format_polynomial(f, prime) for f in F) + "\n")
