"""Shared helpers for the parity tests (golden fixture access, tuple-poly conversions)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "episodes.json")
_cache = {}


def golden():
    if "g" not in _cache:
        with open(GOLDEN) as f:
            _cache["g"] = json.load(f)
    return _cache["g"]


def expand(F):
    """golden compact ideal -> tuple polys with 8 exponent slots"""
    return [[(t[0], tuple(t[1:]) + (0,) * (8 - len(t) + 1)) for t in f] for f in F]


def episode_id(rec):
    how = rec.get("selection", "replay")
    return "%s-s%d-%s-%s%s%s%s" % (rec["dist"], rec["seed"], how, rec["elimination"],
                                   "" if rec["sort_reducers"] else "-unsorted", "-sortin" if rec["sort_input"] else "",
                                   "-red" if rec["rewards"] == "reductions" else "")


def run_on_oracle(orc, rec):
    """Replays a golden episode on an oracle; returns (ideal, pairs0, trace, final_gb, basis)."""
    env = orc.env(rec["dist"], elimination=rec["elimination"], rewards=rec["rewards"],
                  sort_reducers=rec["sort_reducers"], sort_input=rec["sort_input"])
    env.seed(rec["seed"])
    G0, P0 = env.reset()
    if "selection" in rec:
        trace = env.run(selection=rec["selection"])
    else:
        trace = env.run(actions=rec["actions"])
    return G0, P0, np.asarray(trace), env.final_gb(), env.basis(), env


def best_oracle():
    """The unmodified reference when oracle/_ref is built, else the C restatement."""
    from oracle import oracle as O
    return O.load_ref() if O.have_ref() else O.load_port()
