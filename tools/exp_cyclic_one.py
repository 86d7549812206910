"""Runs ONE cyclic-6 episode (Random selection stream seed argv[1]) alone on the GPU: the serial chain of a launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepgroebner_b200.buchberger import BuchbergerEngine
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1234 + 241
eng = BuchbergerEngine("cyclic-6", num_envs=1)
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    stats, _ = eng.run_episodes("random", episodes=1, selection_seed=seed)
    b.record(); torch.cuda.synchronize()
    print("%.1f ms, %d additions, %d steps" % (a.elapsed_time(b), stats["additions"][0], stats["steps"][0]))
