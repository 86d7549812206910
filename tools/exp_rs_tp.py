"""Diagnostic: cyclic-6, EPISODES (default 8192) episodes in one launch on a given number of environment slots:
  python tools/exp_rs_tp.py mode:slots [mode:slots ...]   (mode = bb_set_wide)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
E = int(os.environ.get("EPISODES", "8192"))
from deepgroebner_b200.buchberger import BuchbergerEngine
for spec in sys.argv[1:]:
    mode, slots = [int(x) for x in spec.split(":")]
    eng = BuchbergerEngine("cyclic-6", num_envs=slots)
    eng.set_wide(mode)
    eng.run_episodes("random", episodes=64, selection_seed=1234)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    st, _ = eng.run_episodes("random", episodes=E, selection_seed=1234)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("mode %d, %d slots: %d episodes: %.1f ms; adds/s %.1f M" % (mode, slots, E, ms, st["additions"].sum() / ms / 1e3), flush=True)
    del eng
