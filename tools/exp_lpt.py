"""Diagnostic: how much of k_run's time is queue-order tail?  Runs the 16384-episode workload with seeds in natural
order, then sorted by (measured) episode length descending (perfect LPT) and ascending (worst case).
Historical (v5): since bb_run orders its queue itself (k_order, longest predicted first) the order of `seeds` no longer
matters and all three runs take the same time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepgroebner_b200.buchberger import BuchbergerEngine, resident_envs

slots = int(sys.argv[1]) if len(sys.argv) > 1 else resident_envs(0, 3)
eng = BuchbergerEngine("3-20-10-weighted", num_envs=slots)
E = 16384
def run(seeds, reps=5):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        stats, _ = eng.run_episodes("degree", episodes=E, seeds=seeds, to_host=False)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    import ctypes as C
    from deepgroebner_b200 import _lib
    st = stats.cpu().numpy().view(np.dtype(_lib.STATS_DTYPE))[:E]
    return min(ts), st
seeds = np.arange(E, dtype=np.int32)
t0, st = run(seeds)
steps = st["steps"].astype(np.int64) + 10
print("slots", slots, "natural order: %.3f ms" % t0, "total steps", st["steps"].sum())
for name, order in (("perfect LPT", np.argsort(-steps, kind="stable")), ("shortest first", np.argsort(steps, kind="stable"))):
    t, _ = run(seeds[order].copy())
    print("%s: %.3f ms  (%.2fx)" % (name, t, t0 / t))
