"""Host side of the make_strat CSV path: the ideal-file grammar of scripts/make_strat.cpp /
polynomials.cpp:226-300, checked against the reference's own known answers and against the oracle's parser."""
import numpy as np
import pytest

from deepgroebner_b200.strat import format_polynomial, parse_ideal_string, parse_polynomial, read_ideal_file, write_ideal_file
from helpers import best_oracle

P = 32003


def e8(*v):
    return tuple(v) + (0,) * (8 - len(v))


def canon(f):
    return sorted((c % P, tuple(e)) for c, e in f)


def test_parse_polynomial_reference_known_answers():
    """tests/test_polynomials.cpp:219-245."""
    assert canon(parse_polynomial("a^2*b+c*d")) == canon([(1, e8(2, 1)), (1, e8(0, 0, 1, 1))])
    assert canon(parse_polynomial("413*a^2*b^5*c+32*d^2-5")) == canon([(413, e8(2, 5, 1)), (32, e8(0, 0, 0, 2)), (P - 5, e8())])
    assert canon(parse_polynomial("3")) == [(3, e8())]
    assert canon(parse_polynomial("12*a^2-b*c+13*d")) == canon([(12, e8(2)), (P - 1, e8(0, 1, 1)), (13, e8(0, 0, 0, 1))])


def test_parse_polynomial_sums_equal_monomials_and_rejects_bad_names():
    assert canon(parse_polynomial("a*b+2*b*a-3*a*b")) == []          # cancels to zero
    assert canon(parse_polynomial("a*a^2+5")) == canon([(1, e8(3)), (5, e8())])
    with pytest.raises(ValueError):
        parse_polynomial("a+z")


def test_parser_matches_oracle_on_random_polynomials():
    orc = best_oracle()
    rng = np.random.default_rng(11)
    for _ in range(300):
        f = {}
        for _t in range(int(rng.integers(1, 7))):
            e = tuple(int(x) for x in rng.integers(0, 4, size=5)) + (0, 0, 0)
            f[e] = int(rng.integers(1, P))
        poly = list(f.items())
        s = format_polynomial([(c, e) for e, c in poly])
        assert canon(parse_polynomial(s)) == canon(orc.parse_polynomial(s)) == canon([(c, e) for e, c in poly]), s


def test_ideal_file_round_trip(tmp_path):
    orc = best_oracle()
    gen = orc.generator("3-20-10-weighted")
    gen.seed(5)
    ideals = [gen.next() for _ in range(20)]
    path = str(tmp_path / "data" / "stats" / "d" / "d.csv")
    write_ideal_file(path, ideals)
    back = read_ideal_file(path)
    assert [[canon(f) for f in F] for F in back] == [[canon(f) for f in F] for F in ideals]
    assert open(path).readline() == "Ideal\n"
    assert parse_ideal_string("a^2*b+c*d|b^3-7*a")[1] == [(1, e8(0, 3)), (P - 7, e8(1))]
