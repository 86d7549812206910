"""Host-side recomputation of the device checksums (definition in include/bbenv.h, "checksums")."""
import numpy as np

GOLD = np.uint64(0x9E3779B97F4A7C15)
GOLD2 = np.uint64(0xD1B54A32D192ED03)
M1 = np.uint64(0xbf58476d1ce4e5b9)
M2 = np.uint64(0x94d049bb133111eb)


def mix64(z):
    z = np.asarray(z, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        z ^= z >> np.uint64(30)
        z *= M1
        z ^= z >> np.uint64(27)
        z *= M2
        z ^= z >> np.uint64(31)
    return z


def hash_item(x, pos):
    with np.errstate(over="ignore"):
        return mix64(np.asarray(x, np.uint64) + GOLD * (np.asarray(pos, np.uint64) + np.uint64(1)))


def _sum(a):
    with np.errstate(over="ignore"):
        return int(np.sum(a, dtype=np.uint64))


def trace_hash(trace):
    """trace: int array [T, >=3] with columns i, j, additions"""
    t = np.asarray(trace, dtype=np.uint64).reshape(-1, np.asarray(trace).shape[-1])
    if len(t) == 0:
        return 0
    # h_T = sum_t x_t * GOLD^(T-1-t) mod 2^64  (h <- h * GOLD + x_t, include/bbenv.h "checksums")
    with np.errstate(over="ignore"):
        x = (t[:, 0] | (t[:, 1] << np.uint64(16)) | (t[:, 2] << np.uint64(32))) + np.uint64(1)
        powers = np.cumprod(np.concatenate([np.ones(1, np.uint64), np.full(len(t) - 1, GOLD, np.uint64)]), dtype=np.uint64)[::-1]
        return int(np.sum(x * powers, dtype=np.uint64))


def polys_hash(polys):
    """polys: [[(coef, exps), ...], ...] in order"""
    coefs, elo, ehi, lens = [], [], [], []
    for f in polys:
        lens.append(len(f))
        for c, e in f:
            e = tuple(e) + (0,) * (8 - len(e))
            coefs.append(c)
            elo.append(e[0] | (e[1] << 16) | (e[2] << 32) | (e[3] << 48))
            ehi.append(e[4] | (e[5] << 16) | (e[6] << 32) | (e[7] << 48))
    h = 0
    if coefs:
        t = np.arange(len(coefs), dtype=np.uint64)
        three = np.uint64(3)
        h += _sum(hash_item(np.array(coefs, np.uint64), three * t))
        h += _sum(hash_item(np.array(elo, np.uint64), three * t + np.uint64(1)))
        h += _sum(hash_item(np.array(ehi, np.uint64), three * t + np.uint64(2)))
    if lens:
        with np.errstate(over="ignore"):
            q = np.arange(1, len(lens) + 1, dtype=np.uint64)
            h += _sum(mix64(np.array(lens, np.uint64) + GOLD2 * q))
    return h & ((1 << 64) - 1)
