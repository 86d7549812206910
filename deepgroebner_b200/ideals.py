"""Host-side description of input ideals: the ``ideal_dist`` string grammar and fixed ideals.

Mirrors what the environment constructor needs from the reference's ``parse_ideal_dist``
(deepgroebner/ideals.cpp:103-143, deepgroebner/ideals.py:112-139) and ``FixedIdealGenerator``
(ideals.h:116-138, ideals.py:142-166).  Random ideals are NOT generated here: the spec is handed to the CUDA
library, which draws them on device from per-environment minstd_rand0 streams (bb_set_distribution for binomials,
bb_set_distribution_poly for Poisson-length polynomials).
"""
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

Poly = List[Tuple[int, Tuple[int, ...]]]  # [(coef, exponent vector)], any order


@dataclass
class BinomialSpec:
    """'n-d-s-{uniform,weighted,maximum}[-consts][-homog][-pure]' -> RandomBinomialIdealGenerator (ideals.cpp:157-201)"""
    n: int
    d: int
    s: int
    dist: str
    constants: bool = False
    homogeneous: bool = False
    pure: bool = False


@dataclass
class PolySpec:
    """'n-d-s-lam-{uniform,weighted,maximum}[-consts][-homog]' -> RandomIdealGenerator (ideals.cpp:204-231):
    s monic polynomials of 2 + Poisson(lam) random terms each."""
    n: int
    d: int
    s: int
    lam: float
    dist: str
    constants: bool = False
    homogeneous: bool = False

    def max_gen_terms(self):
        """Term capacity of one staged ideal: every polynomial at 2 + (lam + 10 sqrt(lam) + 10) terms -- a Poisson
        tail of < 1e-9 per polynomial; an ideal beyond it is flagged BB_STATUS_OVERFLOW_TERMS, never truncated."""
        return self.s * (2 + int(self.lam + 10.0 * self.lam ** 0.5 + 10.0))


@dataclass
class FixedIdealGenerator:
    """An explicit list of generators, replayed at every reset (ideals.py:142-166)."""
    F: Sequence[Poly]
    n: int = 0  # number of variables; 0 = infer from the largest variable index present

    def nvars(self):
        if self.n:
            return self.n
        top = 0
        for f in self.F:
            for _, e in f:
                for i, x in enumerate(e):
                    if x:
                        top = max(top, i + 1)
        return max(top, 1)


def cyclic(n: int, prime: int = 32003) -> List[Poly]:
    """The cyclic-n system (ideals.cpp:16-36): for d = 1..n-1 the sum over i of prod_{k<d} x_{(i+k) mod n},
    and x_0...x_{n-1} - 1."""
    F = []
    for d in range(1, n):
        f = []
        for i in range(n):
            e = [0] * n
            for k in range(d):
                e[(i + k) % n] = 1
            f.append((1, tuple(e)))
        F.append(f)
    F.append([(1, tuple([1] * n)), (prime - 1, tuple([0] * n))])
    return F


def parse_ideal_dist(ideal_dist: str, prime: int = 32003):
    """Returns a BinomialSpec or a FixedIdealGenerator for the reference's ideal_dist strings."""
    args = ideal_dist.split("-")
    if args[0] == "cyclic":
        n = int(args[1])
        return FixedIdealGenerator(cyclic(n, prime), n)
    if len(args) >= 4 and args[3] in ("uniform", "weighted", "maximum"):
        return BinomialSpec(int(args[0]), int(args[1]), int(args[2]), args[3], "consts" in args, "homog" in args,
                            "pure" in args)
    if len(args) >= 5 and args[4] in ("uniform", "weighted", "maximum"):
        return PolySpec(int(args[0]), int(args[1]), int(args[2]), float(args[3]), args[4], "consts" in args,
                        "homog" in args)
    raise ValueError("cannot parse ideal_dist %r" % ideal_dist)
