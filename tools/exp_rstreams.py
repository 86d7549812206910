"""Diagnostic: the register-stream reducer (bb_set_wide(4..6), bb_rstreams.cuh) against the CTA-per-environment one (mode 1)
on cyclic-6: records equal, the longest episode alone, one launch of 1024 / 8192 episodes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepgroebner_b200.buchberger import BuchbergerEngine
pmodes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1, 2, 3, 7, 4, 5, 6]   # compared on 48 episodes
modes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4]                    # timed
big = len(sys.argv) > 3 and sys.argv[3] == "big"
eng = BuchbergerEngine("cyclic-6", num_envs=int(os.environ.get("NUM_ENVS", "1024")))
ref = None
for mode in pmodes:
    eng.set_wide(mode)
    st, _ = eng.run_episodes("random", episodes=48, selection_seed=1234, compute_gb=True)
    if ref is None: ref = st
    bad = [f for f in st.dtype.names if not np.array_equal(st[f], ref[f])]
    print("mode %d: 48 episodes, status ok %s, fields differing from mode %d: %s" % (mode, bool((st["status"] == 2).all()), pmodes[0], bad))
one = BuchbergerEngine("cyclic-6", num_envs=1)
for mode in modes:
    one.set_wide(mode)
    for rep in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        st, _ = one.run_episodes("random", episodes=1, selection_seed=1234 + 241)
        b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("mode %d: longest episode alone %.1f ms, %d additions, %.3f us per addition" % (mode, ms, st["additions"][0], ms * 1e3 / st["additions"][0]))
for mode in modes:
    eng.set_wide(mode)
    for n in ((1024, 8192) if big else (1024,)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        st, _ = eng.run_episodes("random", episodes=n, selection_seed=1234)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print("mode %d: episodes %d: %.1f ms; adds/s %.1f M" % (mode, n, ms, st["additions"].sum() / ms / 1e3))
