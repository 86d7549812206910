"""Diagnostic: A/B of k_run_wide builds (BBENV_LIB=deepgroebner_b200/libbbenv_<variant>.so, `make variant`) on cyclic-6:
the longest episode alone and one launch of 1024 episodes; with the BBW_INSTR build, rounds per addition."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepgroebner_b200.buchberger import BuchbergerEngine
tag = os.path.basename(os.environ.get("BBENV_LIB", "libbbenv.so"))
one = BuchbergerEngine("cyclic-6", num_envs=1)
for rep in range(2):
    one.counters(reset=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    st, _ = one.run_episodes("random", episodes=1, selection_seed=1234 + 241)
    b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print("%s: longest episode alone %.1f ms, %.3f us per addition; counters %s" % (tag, ms, ms * 1e3 / st["additions"][0], one.counters()))
eng = BuchbergerEngine("cyclic-6", num_envs=1024)
eng.run_episodes("random", episodes=64, selection_seed=1234)
for n in (1024, 8192) if len(sys.argv) > 1 else (1024,):
    eng.counters(reset=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    st, _ = eng.run_episodes("random", episodes=n, selection_seed=1234)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("%s: episodes %d: %.1f ms; adds/s %.1f M; counters %s" % (tag, n, ms, st["additions"].sum() / ms / 1e3, eng.counters()))
