"""Diagnostic: the serial chain that bounds a cyclic-6 launch.  Runs E episodes, then ONLY the episode with the most
additions (alone on the GPU: its time is the lower bound of any launch that contains it)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepgroebner_b200.buchberger import BuchbergerEngine

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = BuchbergerEngine("cyclic-6", num_envs=E)
def timed(**kw):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    stats, _ = eng.run_episodes("random", **kw)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b), stats
timed(episodes=E, selection_seed=1234)
ms, st = timed(episodes=E, selection_seed=1234)
adds = st["additions"].astype(np.int64)
print("all %d episodes: %.1f ms, %d additions, %d steps" % (E, ms, adds.sum(), st["steps"].sum()))
order = np.argsort(-adds)
for e in order[:3]:
    # episode e alone: same selection stream (seed 1234 + e)
    ms1, s1 = timed(episodes=1, selection_seed=1234 + int(e))
    assert s1["additions"][0] == adds[e]
    print("episode %d alone: %.1f ms, %d additions (%.2f us each), %d steps, basis %d" % (
        e, ms1, adds[e], 1000.0 * ms1 / adds[e], s1["steps"][0], s1["nbasis"][0]))
print("percentiles of additions per episode:", np.percentile(adds, [50, 90, 99, 100]).astype(int))
