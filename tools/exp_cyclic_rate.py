"""Diagnostic: cyclic-6 throughput of one launch at several batch sizes (chain-bound at 1024, throughput-bound at 8192)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepgroebner_b200.buchberger import BuchbergerEngine
eng = BuchbergerEngine("cyclic-6", num_envs=1024)
eng.run_episodes("random", episodes=64, selection_seed=1234)
for n in (1024, 8192):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    st, _ = eng.run_episodes("random", episodes=n, selection_seed=1234)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("episodes %d: %.1f ms; adds/s %.1f M" % (n, ms, st["additions"].sum() / ms / 1e3))
