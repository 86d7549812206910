"""ref_run_records (oracle/ref_shim.cpp) -- the whole-run record generator behind bench.py's exhaustive parity gate and
the full-size GPU tests -- pinned against the reference functions it is built from: the env's own step trace
(ref_env_run), the reference's buchberger() loop for seeded Random selection, and the golden episodes."""
import numpy as np
import pytest

from hashing import polys_hash, trace_hash
from helpers import golden


@pytest.mark.parametrize("dist,strategy", [("3-20-10-weighted", "degree"), ("3-20-10-weighted", "first"),
                                           ("3-20-10-uniform", "normal"), ("5-5-10-uniform", "sugar"),
                                           ("3-6-5-0.5-uniform", "degree")])
def test_records_equal_step_by_step_env(ref, dist, strategy):
    n = 96
    rec = ref.run_records(dist, strategy, n, seed0=40, compute_gb=True, nthreads=3)
    env = ref.env(dist)
    for e in range(n):
        env.seed(40 + e)
        G0, _ = env.reset()
        t = env.run(selection=strategy)
        r = rec[e]
        assert r["status"] == 2 and r["steps"] == len(t) and r["additions"] == int(t[:, 2].sum())
        assert r["nbasis"] == len(env.basis()) and r["nterms"] == sum(len(g) for g in env.basis())
        assert r["nonzero_reductions"] == r["nbasis"] - len(G0) and r["zero_reductions"] == r["steps"] - r["nonzero_reductions"]
        assert int(r["trace_hash"]) == trace_hash(t)
        assert int(r["basis_hash"]) == polys_hash(env.basis())
        gb = env.final_gb()
        assert int(r["gb_hash"]) == polys_hash(gb) and (r["gb_polys"], r["gb_terms"]) == (len(gb), sum(len(g) for g in gb))
        ret, disc = 0.0, 1.0
        for a in t[:, 2]:
            ret += disc * -float(a)
            disc *= 0.99
        assert r["discounted_return"] == ret


def test_records_random_selection_equal_reference_buchberger_loop(ref):
    """Random: choice() on minstd_rand0 seeded sel_seed0 + e * stride, as buchberger(..., seed) (buchberger.cpp:190-197)."""
    n = 24
    rec = ref.run_records("3-20-10-weighted", "random", n, seed0=7, sel_seed0=500, sel_stride=3, gamma=0.9)
    env = ref.env("3-20-10-weighted")
    for e in range(n):
        env.seed(7 + e)
        F, _ = env.reset()
        gb, st = ref.buchberger(F, selection="random", gamma=0.9, seed=500 + 3 * e)
        r = rec[e]
        assert (r["zero_reductions"], r["nonzero_reductions"], r["additions"]) == \
            (st["zero_reductions"], st["nonzero_reductions"], st["polynomial_additions"])
        assert r["discounted_return"] == st["discounted_return"] and int(r["gb_hash"]) == polys_hash(gb)


def test_records_fixed_ideal_truncation_and_thread_independence(ref):
    a = ref.run_records("cyclic-4", "normal", 5, compute_gb=True, nthreads=1)
    b = ref.run_records("cyclic-4", "normal", 5, compute_gb=True, nthreads=4)
    assert a.tobytes() == b.tobytes() and len(set(a["trace_hash"].tolist())) == 1
    t = ref.run_records("3-20-10-weighted", "degree", 8, max_steps=5)
    assert (t["steps"] <= 5).all() and ((t["status"] == 1) | (t["steps"] < 5)).all()
    assert (t["gb_hash"][t["status"] == 1] == 0).all()


def test_records_match_golden_episodes(ref):
    """The committed golden episodes (recorded from the unmodified reference) through the record path."""
    done = 0
    for rec in golden()["episodes"]:
        if "selection" not in rec or rec["rewards"] != "additions" or rec["selection"] not in ("first", "degree", "normal", "sugar"):
            continue
        if rec["dist"].startswith("cyclic") or "-" not in rec["dist"]:
            continue
        r = ref.run_records(rec["dist"], rec["selection"], 1, seed0=rec["seed"], elimination=rec["elimination"],
                            sort_input=rec["sort_input"], sort_reducers=rec["sort_reducers"])[0]
        assert r["steps"] == len(rec["trace"]) and int(r["trace_hash"]) == trace_hash(np.array(rec["trace"]))
        done += 1
    assert done >= 5
