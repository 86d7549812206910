"""Diagnostic: cyclic-n episodes under seeded Random selection with both runners and several dividend capacities."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepgroebner_b200.buchberger import BuchbergerEngine

name = sys.argv[1] if len(sys.argv) > 1 else "cyclic-6"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
for cap, mode in [(int(c), int(m)) for c, m in (x.split(":") for x in (sys.argv[3] if len(sys.argv) > 3 else "1024:1,2048:1").split(","))]:
    eng = BuchbergerEngine(name, num_envs=E, max_poly_terms=cap)
    eng.set_wide(mode)
    ts = []
    for rep in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        stats, _ = eng.run_episodes("random", episodes=E, selection_seed=1234)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    st, cnt = np.unique(stats["status"], return_counts=True)
    ok = stats["status"] == 2
    print("cap %d wide %d: %.1f ms  status %s  steps %d adds %d  max adds/episode %d mean %d" % (
        cap, mode, min(ts), dict(zip(st.tolist(), cnt.tolist())), stats["steps"][ok].sum(), stats["additions"][ok].sum(),
        stats["additions"].max(), stats["additions"].mean()), flush=True)
    del eng
