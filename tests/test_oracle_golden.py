"""The C restatement (oracle/bb_oracle.c) against golden episodes recorded from the UNMODIFIED reference
(tests/golden/episodes.json, made by tests/golden/make_golden.py), plus -- when oracle/_ref is present --
a randomized function-by-function cross-check of the restatement against the reference itself."""
import numpy as np
import pytest

from helpers import episode_id, expand, golden, run_on_oracle

EPISODES = golden()["episodes"]


@pytest.mark.parametrize("rec", EPISODES, ids=[episode_id(r) for r in EPISODES])
def test_port_replays_golden_episode(port, rec):
    G0, P0, trace, gb, basis, env = run_on_oracle(port, rec)
    assert G0 == expand(rec["ideal"])                       # generator stream (minstd_rand0 + libstdc++ dists)
    assert [list(p) for p in P0] == rec["pairs0"]           # reset(): update per generator
    assert trace.tolist() == rec["trace"]                   # pair sequence, additions, |P|, |G| per step
    assert gb == expand(rec["final_gb"])                    # interreduce(minimalize(G))
    assert [len(g) for g in basis] == rec["basis_len"]
    assert basis[-3:] == expand(rec["last_basis"])
    if "value_0.99" in rec:
        env.seed(rec["seed"])
        env.reset()
        assert env.value(rec["selection"], 0.99) == rec["value_0.99"]


@pytest.mark.parametrize("rec", golden()["lm"], ids=lambda r: "%s-k%d" % (r["dist"], r["k"]))
def test_port_lead_monomials_matrices(port, rec):
    env = port.lm_env(rec["dist"], k=rec["k"])
    env.seed(rec["seed"])
    s = env.reset()
    assert s.tolist() == rec["states"][0]
    for a, r, exp in zip(rec["actions"], rec["rewards"], rec["states"][1:]):
        s, reward, done, _ = env.step(a)
        assert reward == r and s.tolist() == exp and done == (len(exp) == 0)


def test_port_value_known_answer(port):
    v = golden()["values"][0]
    env = port.lm_env(v["dist"], k=v["k"])
    env.seed(v["seed"])
    env.reset()
    _, r, _, _ = env.step(v["after_action"])
    assert r == v["reward"] == -1.0
    assert env.value("degree", 0.99) == v["value_degree"] == -123.7032989562525
    assert env.value("first", 0.99) == v["value_first"]
    assert env.value("normal", 0.99) == v["value_normal"]


def test_survey_golden_3_20_10_weighted_seed123(port):
    """SURVEY 8(c) 'extra golden': first ideal, 19 pairs, Degree 131 steps / 248 additions, First 105 / 349."""
    env = port.env("3-20-10-weighted")
    env.seed(123)
    G, P = env.reset()
    assert G[0] == [(1, (0, 14, 5, 0, 0, 0, 0, 0)), (31, (1, 4, 6, 0, 0, 0, 0, 0))]
    assert G[9] == [(1, (4, 3, 9, 0, 0, 0, 0, 0)), (7715, (3, 1, 0, 0, 0, 0, 0, 0))]
    assert len(P) == 19
    t = env.run(selection="degree")
    assert len(t) == 131 and int(t[:, 2].sum()) == 248
    assert t[:5, :4].tolist() == [[2, 6, 1, 17], [2, 10, 1, 17], [6, 10, 1, 17], [2, 12, 1, 16], [10, 12, 1, 17]]
    y, x2z2 = [(1, (0, 1, 0, 0, 0, 0, 0, 0))], [(1, (2, 0, 2, 0, 0, 0, 0, 0))]
    assert env.final_gb() == [y, x2z2]
    env.seed(123)
    env.reset()
    t = env.run(selection="first")
    assert len(t) == 105 and int(t[:, 2].sum()) == 349
    assert t[:6, :4].tolist() == [[1, 2, 1, 20], [2, 3, 1, 16], [0, 4, 1, 17], [0, 5, 1, 17], [2, 6, 1, 18], [0, 7, 6, 19]]
    assert env.final_gb() == [y, x2z2]


# ------------------------------------------------------------------ live cross-check vs the reference (if built)
def _rand_poly(rng, orc, n, nterms, maxe, p=32003):
    terms = {}
    nterms = min(nterms, (maxe + 1) ** n)
    while len(terms) < nterms:
        e = tuple(int(x) for x in rng.integers(0, maxe + 1, n)) + (0,) * (8 - n)
        terms[e] = int(rng.integers(1, p))
    return orc.poly_make([(c, e) for e, c in terms.items()])


def test_port_vs_reference_primitives(port, ref):
    rng = np.random.default_rng(11)
    for it in range(300):
        n = int(rng.integers(1, 9))
        f = _rand_poly(rng, ref, n, int(rng.integers(1, 7)), 4)
        g = _rand_poly(rng, ref, n, int(rng.integers(1, 7)), 4)
        assert port.poly_make(f) == f
        assert port.poly_add(f, g) == ref.poly_add(f, g)
        assert port.poly_sub(f, g) == ref.poly_sub(f, g)
        assert port.spoly(f, g) == ref.spoly(f, g)
        assert port.mono_cmp(f[0][1], g[0][1]) == ref.mono_cmp(f[0][1], g[0][1])
        F = [_rand_poly(rng, ref, n, int(rng.integers(1, 5)), 3) for _ in range(int(rng.integers(1, 6)))]
        assert port.reduce(f, F) == ref.reduce(f, F)
        pairs = [(i, j) for j in range(len(F)) for i in range(j) if rng.random() < 0.6]
        for e in ("gebauermoeller", "lcm", "none"):
            assert port.update(F, pairs, f, e) == ref.update(F, pairs, f, e)
        assert port.minimalize(F) == ref.minimalize(F)
    for a, b in rng.integers(1, 32003, (500, 2)):
        assert port.c_coef_div(int(a), int(b)) == ref.c_coef_div(int(a), int(b))


@pytest.mark.parametrize("dist", ["3-20-10-weighted", "3-20-10-uniform", "5-5-10-uniform", "4-8-6-0.5-weighted"])
def test_port_vs_reference_episodes(port, ref, dist):
    for seed in range(100, 112):
        for sel in ("degree", "first", "normal", "sugar"):
            out = []
            for o in (port, ref):
                env = o.env(dist)
                env.seed(seed)
                G0, P0 = env.reset()
                v = env.value(sel, 0.99)
                t = env.run(selection=sel)
                out.append((G0, P0, v, t.tolist(), env.final_gb(), env.basis()))
            assert out[0] == out[1]


@pytest.mark.parametrize("dist", ["3-20-10-weighted", "5-5-10-uniform", "cyclic-4"])
def test_port_value_and_all_strategies_equal_reference(port, ref, dist):
    """The seeded value() of both oracles (all nine SelectionTypes, best-of-random, 'sample') agree bit for bit, and
    the unseeded deterministic strategies equal BuchbergerEnv::value itself."""
    strategies = ["first", "degree", "normal", "sugar", "random", "last", "codegree", "strange", "spice"]
    for seed in (1, 2, 3):
        a, b = port.env(dist), ref.env(dist)
        for env in (a, b):
            env.seed(seed)
            env.reset()
        for pick in (0, 1, 0):
            P = a.pairs()
            if not P:
                break
            assert P == b.pairs()
            for s in strategies:
                assert a.value_seeded(s, 0.99, 9) == b.value_seeded(s, 0.99, 9), (dist, seed, s)
            for s in ("first", "degree", "normal", "sugar"):
                assert a.value_seeded(s, 0.9) == b.value(s, 0.9)
            assert a.value_seeded("random", 0.99, 4, 6) == b.value_seeded("random", 0.99, 4, 6)
            assert a.value_seeded("sample", 0.99, 2, 12) == b.value_seeded("sample", 0.99, 2, 12)
            act = P[min(pick, len(P) - 1)]
            assert a.step(act) == b.step(act)
