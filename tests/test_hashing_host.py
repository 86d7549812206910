"""CPU tests of tests/hashing.py, the host-side mirror of the device checksums (include/bbenv.h, "checksums"): the
vectorised forms equal the step-by-step definitions, so a GPU parity failure points at the device, not at the checker."""
import numpy as np

from hashing import GOLD, hash_item, mix64, polys_hash, trace_hash

M64 = (1 << 64) - 1


def fold(trace):
    h = 0
    for i, j, a in trace:
        h = (h * int(GOLD) + ((i | (j << 16) | (a << 32)) + 1)) & M64
    return h


def test_trace_hash_is_the_rolling_hash_of_the_header():
    rng = np.random.default_rng(0)
    for T in (0, 1, 2, 7, 257, 5000):
        tr = np.stack([rng.integers(0, 500, T), rng.integers(0, 500, T), rng.integers(1, 4000, T)], axis=1)
        assert trace_hash(tr) == fold(tr.tolist())
    # order-sensitive, and extra columns (|P| after, |G|) are ignored
    tr = np.array([[0, 3, 1, 9, 4], [2, 5, 7, 8, 4], [1, 4, 2, 7, 5]])
    assert trace_hash(tr) == fold(tr[:, :3].tolist()) != trace_hash(tr[::-1])


def test_hash_item_is_splitmix64_of_the_salted_value():
    def scalar(x, pos):
        z = (x + int(GOLD) * (pos + 1)) & M64
        z ^= z >> 30; z = (z * 0xbf58476d1ce4e5b9) & M64
        z ^= z >> 27; z = (z * 0x94d049bb133111eb) & M64
        return z ^ (z >> 31)
    xs = np.array([0, 1, 32003, 2 ** 40 + 17, M64], dtype=np.uint64)
    ps = np.array([0, 1, 5, 99, 2 ** 33], dtype=np.uint64)
    got = hash_item(xs, ps)
    assert [int(g) for g in got] == [scalar(int(x), int(p)) for x, p in zip(xs, ps)]
    assert int(mix64(np.uint64(0))) == 0


def test_polys_hash_depends_on_order_lengths_and_padding():
    f = [(1, (2, 0, 1)), (5, (0, 1, 0))]
    g = [(1, (1, 1, 0)), (32002, (0, 0, 3))]
    assert polys_hash([f, g]) != polys_hash([g, f])
    assert polys_hash([f + g]) != polys_hash([f, g])                      # same terms, different polynomial lengths
    assert polys_hash([f]) == polys_hash([[(c, e + (0, 0)) for c, e in f]])  # exponent vectors are zero-padded to 8
    assert polys_hash([]) == 0
