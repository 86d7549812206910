"""Diagnosis: which cyclic-6 episodes of a large seeded-Random batch end in a fault status, per runner."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from deepgroebner_b200.buchberger import BuchbergerEngine
E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
eng = BuchbergerEngine("cyclic-6", num_envs=1024)
for mode in (1, 2, 0):
    eng.set_wide(mode)
    st, _ = eng.run_episodes("random", episodes=E, selection_seed=1234)
    bad = np.nonzero(st["status"] != 2)[0]
    print("mode", mode, [(int(e), int(st["status"][e]), int(st["steps"][e]), int(st["additions"][e]), int(st["nbasis"][e]), int(st["nterms"][e])) for e in bad])
