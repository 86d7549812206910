"""Diagnostic: throughput regime of the cyclic-6 runners: 8192 episodes in one launch on NUM_ENVS environment slots, modes argv[1]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepgroebner_b200.buchberger import BuchbergerEngine
for spec in sys.argv[1:]:
    mode, slots = [int(x) for x in spec.split(":")]
    eng = BuchbergerEngine("cyclic-6", num_envs=slots)
    eng.set_wide(mode)
    eng.run_episodes("random", episodes=64, selection_seed=1234)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    st, _ = eng.run_episodes("random", episodes=8192, selection_seed=1234)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("mode %d, %d slots: 8192 episodes: %.1f ms; adds/s %.1f M" % (mode, slots, ms, st["additions"].sum() / ms / 1e3), flush=True)
    del eng
