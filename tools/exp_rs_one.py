"""Diagnostic for ncu: `episodes` cyclic-6 episodes (seeded Random) on as many environments, bb_set_wide(mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepgroebner_b200.buchberger import BuchbergerEngine
mode, n = int(sys.argv[1]), int(sys.argv[2])
eng = BuchbergerEngine("cyclic-6", num_envs=n)
eng.set_wide(mode)
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    st, _ = eng.run_episodes("random", episodes=n, selection_seed=1234)
    b.record(); torch.cuda.synchronize()
    print("%.1f ms, %d additions" % (a.elapsed_time(b), st["additions"].sum()))
