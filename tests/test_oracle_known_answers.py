"""Pins the oracle against the known answers held by the reference's OWN tests for this path
(tests/test_polynomials.cpp, tests/test_ideals.cpp, tests/test_buchberger.cpp, tests/test_buchberger.py).
Each case cites the reference test it restates.  Runs on oracle/bb_oracle.c and (when built) oracle/_ref."""
import numpy as np
import pytest

P = 32003


def T(c, e):
    return (c % P, tuple(e) + (0,) * (8 - len(e)))


def poly(orc, *terms):
    return orc.poly_make([T(c, e) for c, e in terms])


# ---------------------------------------------------------------- tests/test_polynomials.cpp
def test_coefficient(orc):  # :5-49
    o = orc
    assert o.c_coef_norm(2045) == 2045 and o.c_coef_norm(-2) == 32001 and o.c_coef_norm(32008) == 5
    assert o.c_coef_add(3, 10) == 13 and o.c_coef_sub(10, 3) == 7
    assert o.c_coef_mul(3, 10) == 30 and o.c_coef_mul(3, -2) == 31997
    assert o.c_coef_div(3, 10) == 28803
    assert o.c_coef_div(23002, 32001) == 20502
    assert o.c_coef_div(12000, 4) == 3000
    assert o.c_coef_div(12345, 1) == 12345
    assert o.c_coef_div(3, 7) == 13716  # Term divide :139-144


def test_monomial_order(orc):  # :76-86
    m1, m2, m3 = (1,) * 8, (0,) * 7 + (9,), (0, 0, 0, 0, 2, 2, 2, 2)
    assert orc.mono_cmp(m3, m1) == -1 and orc.mono_cmp(m1, m2) == -1 and orc.mono_cmp(m3, m2) == -1
    assert orc.mono_cmp(m1, m3) == 1 and orc.mono_cmp(m2, m1) == 1 and orc.mono_cmp(m2, m3) == 1


def test_monomial_divisible_lcm(orc):  # :88-122
    m1, m2, m3 = (1, 2, 3, 4, 5, 6, 7, 8), (0, 1, 2, 3, 4, 5, 6, 7), (1,) * 8
    exp = [[True, True, True], [False, True, False], [False, False, True]]
    for a, ma in enumerate((m1, m2, m3)):
        for b, mb in enumerate((m1, m2, m3)):
            assert orc.mono_divisible(ma, mb) == exp[a][b]
    a, b = (2, 2, 2, 2, 0, 0, 0, 0), (0, 0, 0, 0, 2, 2, 2, 2)
    assert orc.mono_lcm(a, b) == (2,) * 8 and orc.mono_lcm(b, a) == (2,) * 8
    assert orc.mono_lcm((2,) * 8, (1,) * 8) == (2,) * 8


def test_polynomial_ctor_add_sub(orc):  # :146-186
    p = poly(orc, (1, (1,) * 8), (3, (0, 0, 0, 0, 1, 1, 1, 1)), (9, (1, 1, 2, 2, 3, 4, 1, 1)), (1, ()))
    assert p[0] == T(9, (1, 1, 2, 2, 3, 4, 1, 1))
    p1 = poly(orc, (1, (1, 2, 1)), (3, (1, 0, 1)), (7, ()))
    p2 = poly(orc, (9, (7, 0, 0)), (-3, (1, 0, 1)), (1, (1, 0, 0)))
    p3 = poly(orc, (9, (7, 0, 0)), (1, (1, 2, 1)), (1, (1, 0, 0)), (7, ()))
    assert orc.poly_add(p1, p2) == p3
    assert orc.poly_sub(p3, p2) == p1 and orc.poly_sub(p3, p1) == p2 and orc.poly_sub(p1, p1) == []


def test_polynomial_multiply(orc):  # :188-217
    p1 = poly(orc, (1, (1,) * 8), (3, (0, 0, 0, 0, 1, 1, 1, 1)), (9, (1, 1, 2, 2, 3, 4, 1, 1)), (1, ()))
    p2 = poly(orc, (9, (1, 1, 1, 1, 3, 3, 1, 1)), (27, (0, 0, 0, 0, 3, 3, 1, 1)), (81, (1, 1, 2, 2, 5, 6, 1, 1)),
              (9, (0, 0, 0, 0, 2, 2, 0, 0)))
    t1, t2 = T(9, (0, 0, 0, 0, 2, 2, 0, 0)), T(2, ())
    assert orc.term_mul(t1, p1) == p2 and orc.term_mul(t1, p1)[0][0] == 81
    assert orc.term_mul(t2, p1) == orc.poly_add(p1, p1)
    q1 = poly(orc, (1, (1, 2, 0)), (1, (0, 1, 1)), (1, ()))
    q2 = poly(orc, (1, (1, 1, 1)), (1, (1, 0, 0)))
    q3 = poly(orc, (1, (2, 3, 1)), (1, (1, 2, 2)), (1, (2, 2, 0)), (2, (1, 1, 1)), (1, (1, 0, 0)))
    assert orc.poly_mul(q1, q2) == q3


def test_parse_polynomial(orc):  # :219-246
    assert orc.parse_polynomial("a^2*b+c*d") == poly(orc, (1, (2, 1, 0, 0)), (1, (0, 0, 1, 1)))
    assert orc.parse_polynomial("413*a^2*b^5*c+32*d^2-5") == poly(orc, (413, (2, 5, 1, 0)), (32, (0, 0, 0, 2)), (-5, ()))
    assert orc.parse_polynomial("3") == poly(orc, (3, ()))
    assert orc.parse_polynomial("12*a^2-b*c+13*d") == poly(orc, (12, (2,)), (-1, (0, 1, 1)), (13, (0, 0, 0, 1)))


# ---------------------------------------------------------------- tests/test_ideals.cpp
def test_cyclic3(orc):  # :9-16
    F = orc.cyclic(3)
    assert F == [poly(orc, (1, (1,)), (1, (0, 1)), (1, (0, 0, 1))),
                 poly(orc, (1, (1, 1)), (1, (0, 1, 1)), (1, (1, 0, 1))),
                 poly(orc, (1, (1, 1, 1)), (-1, ()))]


def test_basis(orc):  # :19-40
    assert orc.basis(3, 0) == [(0,) * 8]
    assert [b[:4] for b in orc.basis(4, 1)] == [(1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)]
    assert [b[:3] for b in orc.basis(3, 2)] == [(2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2)]
    assert [b[:3] for b in orc.basis(3, 3)] == [(3, 0, 0), (2, 1, 0), (2, 0, 1), (1, 2, 0), (1, 1, 1), (1, 0, 2),
                                                 (0, 3, 0), (0, 2, 1), (0, 1, 2), (0, 0, 3)]


@pytest.mark.parametrize("n,d,dist,consts,D", [  # :53-120
    (3, 1, "weighted", False, [0.0, 1.0]),
    (3, 1, "weighted", True, [0.5, 0.5]),
    (3, 1, "uniform", True, [0.25, 0.75]),
    (3, 5, "weighted", False, [0.0, 0.2, 0.2, 0.2, 0.2, 0.2]),
    (3, 5, "weighted", True, [1.0 / 6] * 6),
    (3, 5, "uniform", True, [1.0 / 56, 3.0 / 56, 6.0 / 56, 10.0 / 56, 15.0 / 56, 21.0 / 56]),
    (3, 3, "maximum", True, [0.5, 0.0, 0.0, 0.5]),
    (3, 3, "maximum", False, [0.0, 0.0, 0.0, 1.0]),
    (3, 3, "uniform", False, [0.0, 3.0 / 19, 6.0 / 19, 10.0 / 19]),
    (3, 3, "weighted", False, [0.0, 1.0 / 3, 1.0 / 3, 1.0 / 3]),
])
def test_degree_distribution(orc, n, d, dist, consts, D):
    assert orc.degree_distribution(n, d, dist, consts) == D


def test_random_binomial_seed123(orc):  # :123-132 -- pins minstd_rand0 + libstdc++ distributions
    g = orc.generator("3-5-5-uniform")
    g.seed(123)
    F = [poly(orc, (1, (0, 1, 4)), (31, (0, 3, 1))), poly(orc, (1, (3, 1, 1)), (16013, (3, 0, 2))),
         poly(orc, (1, (2, 2, 0)), (18427, (1, 0, 1))), poly(orc, (1, (2, 0, 3)), (15139, (2, 1, 1))),
         poly(orc, (1, (0, 3, 2)), (5374, (1, 0, 2)))]
    assert g.next() == F


def test_random_ideal_seed123(orc):  # :135-144 (Poisson term counts)
    g = orc.generator("3-5-5-0.5-uniform")
    g.seed(123)
    F = [poly(orc, (1, (0, 1, 3)), (22264, (0, 0, 4))), poly(orc, (1, (1, 1, 1)), (1541, (0, 0, 2))),
         poly(orc, (1, (2, 2, 1)), (15981, (0, 2, 1)), (7023, (0, 0, 1))),
         poly(orc, (1, (1, 4, 0)), (10365, (0, 5, 0)), (5289, (1, 3, 0)), (13942, (1, 1, 0))),
         poly(orc, (1, (3, 1, 0)), (11636, (1, 1, 0)))]
    assert g.next() == F


# ---------------------------------------------------------------- tests/test_buchberger.cpp
def test_spoly(orc):  # :9-55
    f = poly(orc, (1, (1, 2, 1)), (3, (1, 0, 1)), (7, ()))
    g = poly(orc, (9, (7, 0, 0)), (-3, (1, 0, 1)), (1, (1, 0, 0)))
    s = poly(orc, (3, (7, 0, 1)), (7, (6, 0, 0)), (10668, (1, 2, 2)), (28447, (1, 2, 1)))
    assert orc.spoly(f, g) == s
    assert orc.spoly(poly(orc, (1, (2, 0)), (1, (1, 1))), poly(orc, (1, (0, 2)), (1, (1, 1)))) == []
    assert orc.spoly(poly(orc, (1, (3, 2)), (-1, (2, 3))), poly(orc, (1, (4, 1)), (1, (0, 2)))) == \
        poly(orc, (-1, (3, 3)), (-1, (0, 3)))
    assert orc.spoly(poly(orc, (1, (2, 0)), (1, (0, 3))), poly(orc, (1, (1, 2)), (1, (1, 0)), (1, ()))) == \
        poly(orc, (1, (3, 0)), (-1, (1, 1)), (-1, (0, 1)))


def test_reduce(orc):  # :58-77
    g = poly(orc, (1, (3, 1, 2)), (1, (2, 0, 1)))
    F = [poly(orc, (1, (2, 0, 0)), (1, (0, 1, 0))), poly(orc, (1, (1, 1, 1)), (1, (0, 0, 1))),
         poly(orc, (1, (1, 0, 2)), (1, (0, 2, 0)))]
    assert orc.reduce(g, F)[0] == poly(orc, (1, (0, 1, 2)), (-1, (0, 1, 1)))
    g = poly(orc, (1, (5, 10, 4)), (22982, (3, 1, 2)))
    F = [poly(orc, (1, (5, 12, 0)), (25797, (1, 5, 2))), poly(orc, (1, (1, 3, 1)), (27630, (2, 1, 0))),
         poly(orc, (1, (1, 9, 1)), (8749, (2, 0, 0)))]
    r, steps = orc.reduce(g, F)
    assert r == poly(orc, (2065, (9, 2, 0)), (22982, (3, 1, 2))) and steps == 4


def test_update_empty(orc):  # :80-102
    f = poly(orc, (1, (2, 0)), (1, (1, 1)), (2, ()))
    for e in ("none", "lcm", "gebauermoeller"):
        assert orc.update([], [], f, e) == []


def test_minimalize_interreduce(orc):  # :105-133
    G = [poly(orc, (1, (1, 2, 0)), (1, (0, 0, 1))), poly(orc, (1, (1, 0, 1)), (3, (0, 1, 0))),
         poly(orc, (1, (2, 0, 0)), (1, (0, 1, 1))), poly(orc, (-3, (0, 3, 0)), (1, (0, 2, 0))),
         poly(orc, (-9, (0, 1, 0)), (-1, (0, 0, 3))), poly(orc, (1, (0, 0, 8)), (243, (0, 0, 1)))]
    Gmin = [poly(orc, (1, (1, 0, 1)), (3, (0, 1, 0))), poly(orc, (1, (2, 0, 0)), (1, (0, 1, 1))),
            poly(orc, (-1, (0, 0, 3)), (-9, (0, 1, 0))), poly(orc, (-3, (0, 3, 0)), (1, (0, 2, 0))),
            poly(orc, (1, (1, 2, 0)), (1, (0, 0, 1)))]
    assert orc.minimalize(G) == Gmin
    Gred = [poly(orc, (1, (1, 0, 1)), (3, (0, 1, 0))), poly(orc, (1, (2, 0, 0)), (1, (0, 1, 1))),
            poly(orc, (1, (0, 0, 3)), (9, (0, 1, 0))), poly(orc, (1, (0, 3, 0)), (21335, (0, 2, 0))),
            poly(orc, (1, (1, 2, 0)), (1, (0, 0, 1)))]
    assert orc.interreduce(Gmin) == Gred


# ---------------------------------------------------------------- tests/test_buchberger.py (GF(32003)/grevlex rows)
def test_py_spoly_reduce_R1(orc):  # :16-18, :31-34 (x,y,z)
    x2xy = poly(orc, (1, (2,)), (1, (1, 1)))
    y2xy = poly(orc, (1, (0, 2)), (1, (1, 1)))
    assert orc.spoly(x2xy, y2xy) == []
    assert orc.spoly(poly(orc, (1, (3, 2)), (-1, (2, 3))), poly(orc, (1, (4, 1)), (1, (0, 2)))) == \
        poly(orc, (-1, (3, 3)), (-1, (0, 3)))


def test_py_update_1(orc):  # :116-123  G=[x*y^2+2xz-x], f=z^5+2x^2yz+xz
    G = [poly(orc, (1, (1, 2, 0)), (2, (1, 0, 1)), (-1, (1,)))]
    f = poly(orc, (1, (0, 0, 5)), (2, (2, 1, 1)), (1, (1, 0, 1)))
    assert orc.update(G, [], f, "none") == [(0, 1)]
    assert orc.update(G, [], f, "lcm") == []
    assert orc.update(G, [], f, "gebauermoeller") == []


def test_py_update_5(orc):  # :162-171 (x,y,z), grevlex
    G = [poly(orc, (1, (1, 2, 0)), (2, (0, 0, 1))), poly(orc, (1, (1, 0, 2)), (-1, (0, 2, 0)), (-1, (0, 0, 1))),
         poly(orc, (1, (1,)), (3, ()))]
    f = poly(orc, (1, (0, 2, 3)), (-1, (0, 2, 0)), (4, (0, 0, 4)), (1, (0, 0, 2)))
    assert orc.update(G, [(0, 2)], f, "none") == [(0, 2), (0, 3), (1, 3), (2, 3)]
    assert orc.update(G, [(0, 2)], f, "lcm") == [(0, 2), (0, 3), (1, 3)]
    assert orc.update(G, [(0, 2)], f, "gebauermoeller") == [(0, 2)]


def test_py_buchberger_gbs(orc):  # :236, :239 twisted cubic and cyclic-3 under all eliminations
    F = [poly(orc, (1, (0, 1)), (-1, (2,))), poly(orc, (1, (0, 0, 1)), (-1, (3,)))]
    G = [poly(orc, (1, (0, 2)), (-1, (1, 0, 1))), poly(orc, (1, (1, 1)), (-1, (0, 0, 1))),
         poly(orc, (1, (2,)), (-1, (0, 1)))]
    F3 = orc.cyclic(3)
    G3 = [poly(orc, (1, (1,)), (1, (0, 1)), (1, (0, 0, 1))), poly(orc, (1, (0, 2)), (1, (0, 1, 1)), (1, (0, 0, 2))),
          poly(orc, (1, (0, 0, 3)), (-1, ()))]
    for e in ("none", "lcm", "gebauermoeller"):
        assert orc.buchberger(F, elimination=e)[0] == G
        assert orc.buchberger(F3, elimination=e)[0] == G3


def _episode_return(env, selection):
    env.reset()
    return -len(env.run(selection=selection))  # rewards='reductions' => -1 per step


@pytest.mark.parametrize("sel", ["first", "degree", "normal"])
def test_py_episode_0(orc, sel):  # :270-281 katsura-style 5 variables, rewards='reductions' -> -28
    a, b, c, d, e = [tuple(1 if i == k else 0 for i in range(5)) for k in range(5)]
    def m(*vs):
        return tuple(sum(v[i] for v in vs) for i in range(5))
    F = [poly(orc, (1, a), (2, b), (2, c), (2, d), (2, e), (-1, ())),
         poly(orc, (1, m(a, a)), (2, m(b, b)), (2, m(c, c)), (2, m(d, d)), (2, m(e, e)), (-1, a)),
         poly(orc, (2, m(a, b)), (2, m(b, c)), (2, m(c, d)), (2, m(d, e)), (-1, b)),
         poly(orc, (1, m(b, b)), (2, m(a, c)), (2, m(b, d)), (2, m(c, e)), (-1, c)),
         poly(orc, (2, m(b, c)), (2, m(a, d)), (2, m(b, e)), (-1, d))]
    env = orc.env("cyclic-3", rewards="reductions")
    env.set_ideal(F)
    assert _episode_return(env, sel) == -28


@pytest.mark.parametrize("elim,ret", [("none", -45), ("lcm", -35), ("gebauermoeller", -11)])
def test_py_episode_1(orc, elim, ret):  # :284-296 cyclic-4 under Normal selection
    env = orc.env("cyclic-4", elimination=elim, rewards="reductions")
    assert _episode_return(env, "normal") == ret


def test_py_buchberger_env_sort_reducers(orc):  # :246-256, grevlex over GF(32003) with a,b,c,d
    # F = [a^2 b d - c^2, a d - b c^2 - d, a - c]; step (0,1)
    F = [poly(orc, (1, (2, 1, 0, 1)), (-1, (0, 0, 2, 0))), poly(orc, (1, (1, 0, 0, 1)), (-1, (0, 1, 2, 0)), (-1, (0, 0, 0, 1))),
         poly(orc, (1, (1,)), (-1, (0, 0, 1)))]
    # NOTE the reference test uses ring R2 (QQ, lex); only the control flow (sorted vs insertion-order reducers
    # giving different remainders) is restated here on the grevlex/GF(p) path: the two remainders must differ.
    out = []
    for sr in (True, False):
        env = orc.env("cyclic-3", sort_reducers=sr)
        env.set_ideal(F)
        env.reset()
        env.step((0, 1))
        out.append(env.basis()[-1] if len(env.basis()) == 4 else None)
    assert out[0] is not None and out[1] is not None


def test_lead_monomials_env_gm(orc):  # :344-362 control flow of LeadMonomialsEnv (GF(32003) instead of FF(101))
    F = [poly(orc, (1, (0, 1)), (-1, (2,))), poly(orc, (1, (0, 0, 1)), (-1, (3,)))]
    env = orc.lm_env("cyclic-4", k=1)
    env.set_ideal(F, nvars=3)
    s = env.reset()
    assert np.array_equal(s, [[2, 0, 0, 3, 0, 0]])
    s, _, done, _ = env.step(0)
    assert np.array_equal(s, [[2, 0, 0, 1, 1, 0]]) and not done
    s, _, done, _ = env.step(0)
    assert np.array_equal(s, [[1, 1, 0, 0, 2, 0]]) and not done
    s, _, done, _ = env.step(0)
    assert done
