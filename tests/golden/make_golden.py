"""Generates tests/golden/episodes.json from the UNMODIFIED reference (oracle/_ref/libdgref.so, built by
`make -C oracle ref` from /root/reference/deepgroebner/*.cpp + oracle/ref_shim.cpp).

Run here (where /root/reference exists):   python tests/golden/make_golden.py
The fixture is committed; the GPU box and CI only read it.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def compact(F, n):
    """ideal -> [[ [coef, e0..e(n-1)], ...], ...] (drops the unused exponent slots)"""
    return [[[c] + list(e[:n]) for c, e in f] for f in F]


def episode(ref, dist, seed, n, selection=None, actions_rng=None, elimination="gebauermoeller", rewards="additions",
            sort_reducers=True, sort_input=False):
    env = ref.env(dist, elimination=elimination, rewards=rewards, sort_reducers=sort_reducers, sort_input=sort_input)
    env.seed(seed)
    G0, P0 = env.reset()
    rec = dict(dist=dist, seed=seed, nvars=n, elimination=elimination, rewards=rewards, sort_reducers=sort_reducers,
               sort_input=sort_input, ideal=compact(G0, n), pairs0=[list(p) for p in P0])
    if selection is not None:
        rec["selection"] = selection
        if selection in ("first", "degree", "normal", "sugar"):
            rec["value_0.99"] = env.value(selection, 0.99)
        trace = env.run(selection=selection)
    else:
        acts = []
        rows = []
        while True:
            P = env.pairs()
            if not P:
                break
            a = int(actions_rng.integers(len(P)))
            acts.append(a)
            r, _ = env.step(P[a])
            rows.append([P[a][0], P[a][1], int(-r), len(env.pairs()), len(env.basis())])
        rec["actions"] = acts
        trace = np.array(rows, dtype=np.int32).reshape(-1, 5)
    rec["trace"] = trace.tolist()  # rows: i, j, additions (=-reward), |P| after, |G| after
    rec["final_gb"] = compact(env.final_gb(), n)
    rec["basis_len"] = [len(g) for g in env.basis()]
    rec["last_basis"] = compact(env.basis()[-3:], n)
    return rec


def main():
    ref = O.load_ref()
    out = dict(meta=dict(source="unmodified reference @94f3183e via oracle/ref_shim.cpp", prime=ref.prime(),
                         trace_columns=["i", "j", "additions", "npairs_after", "nbasis_after"]),
               episodes=[], lm=[], values=[])
    rng = np.random.default_rng(7)
    for seed in [123, 0, 1, 2, 3, 4, 5]:
        for sel in ("first", "degree", "normal"):
            out["episodes"].append(episode(ref, "3-20-10-weighted", seed, 3, selection=sel))
    for seed in [123, 0, 1]:
        out["episodes"].append(episode(ref, "3-20-10-weighted", seed, 3, selection="sugar"))
        out["episodes"].append(episode(ref, "3-20-10-weighted", seed, 3, actions_rng=rng))
    for dist, n in (("3-20-10-uniform", 3), ("5-5-10-uniform", 5)):
        for seed in [123, 0, 1]:
            out["episodes"].append(episode(ref, dist, seed, n, selection="degree"))
        out["episodes"].append(episode(ref, dist, 2, n, selection="first"))
        out["episodes"].append(episode(ref, dist, 3, n, selection="normal"))
        out["episodes"].append(episode(ref, dist, 4, n, actions_rng=rng))
    for elim in ("lcm", "none"):
        for seed in [0, 1]:
            out["episodes"].append(episode(ref, "3-10-5-weighted", seed, 3, selection="degree", elimination=elim))
    out["episodes"].append(episode(ref, "3-20-10-weighted", 6, 3, selection="degree", sort_reducers=False))
    out["episodes"].append(episode(ref, "3-20-10-weighted", 7, 3, selection="degree", sort_input=True))
    out["episodes"].append(episode(ref, "3-20-10-weighted", 8, 3, selection="degree", rewards="reductions"))
    out["episodes"].append(episode(ref, "3-6-5-0.5-uniform", 0, 3, selection="degree"))
    out["episodes"].append(episode(ref, "4-5-6-maximum-homog", 1, 4, selection="normal"))
    for nc in (4, 5):
        for sel in ("first", "degree", "normal"):
            out["episodes"].append(episode(ref, "cyclic-%d" % nc, 0, nc, selection=sel))
        for elim in ("lcm", "none"):
            out["episodes"].append(episode(ref, "cyclic-%d" % nc, 0, nc, selection="normal", elimination=elim,
                                           rewards="reductions"))
    out["episodes"].append(episode(ref, "cyclic-5", 0, 5, actions_rng=rng))
    out["episodes"].append(episode(ref, "cyclic-6", 0, 6, selection="degree"))

    # LeadMonomialsEnv matrices exactly as wrapped.pyx returns them (k=2 and k=1), a few steps each
    for dist, k, seed in (("3-20-10-weighted", 2, 123), ("3-20-10-weighted", 1, 5), ("5-5-10-uniform", 2, 1)):
        env = ref.lm_env(dist, k=k)
        env.seed(seed)
        s = env.reset()
        rec = dict(dist=dist, k=k, seed=seed, states=[s.tolist()], actions=[], rewards=[])
        for _ in range(12):
            if len(s) == 0:
                break
            a = int(rng.integers(len(s)))
            s, r, done, _ = env.step(a)
            rec["actions"].append(a)
            rec["rewards"].append(r)
            rec["states"].append(s.tolist())
        out["lm"].append(rec)

    # value('degree', .99) after reset, step(3) on 3-20-10-weighted seed 123
    env = ref.lm_env("3-20-10-weighted", k=2)
    env.seed(123)
    env.reset()
    _, r, _, _ = env.step(3)
    out["values"].append(dict(dist="3-20-10-weighted", seed=123, k=2, after_action=3, reward=r,
                              value_degree=env.value("degree", 0.99), value_first=env.value("first", 0.99),
                              value_normal=env.value("normal", 0.99)))

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "episodes.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes;", len(out["episodes"]), "episodes")


if __name__ == "__main__":
    main()
