"""Diagnostic: the fused rollout (one kernel, T steps inside) against the same work as two kernels per step (policy head,
then step with auto-reset) -- is the fused kernel's instruction-fetch stall worth more than 2 T launches?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepgroebner_b200 import LeadMonomialsEnv
from deepgroebner_b200.rollout import PairsPolicy

N, T = 16384, 128
env = LeadMonomialsEnv("3-20-10-weighted", k=2, num_envs=N, pmax=64)
eng = env.engine
env.seed(np.arange(N)); eng.set_auto_reset(True); eng.reset()
net = PairsPolicy(eng.cols, 128, torch_seed=0, seed=1, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)
out = eng.rollout(net, T)
for rep in range(2):
    eng.counters(reset=True)
    a, b = ev(), ev(); a.record(); eng.rollout(net, T, counter0=(rep + 1) * T, out=out); b.record(); torch.cuda.synchronize()
    c = eng.counters(reset=True)
    print("fused: %.2f ms, %.1f M env-steps/s" % (a.elapsed_time(b), c["env_steps"] / a.elapsed_time(b) / 1e3))
rew = torch.empty(N, dtype=torch.float64, device="cuda"); done = torch.empty(N, dtype=torch.uint8, device="cuda")
for rep in range(2):
    eng.counters(reset=True)
    a, m, b = ev(), ev(), ev()
    tp = ts = 0.0
    a.record()
    for t in range(T):
        acts, logp = eng.policy(net, counter=1000 * (rep + 1) + t)
        eng.step(acts, reward=rew, done=done)
    b.record(); torch.cuda.synchronize()
    c = eng.counters(reset=True)
    print("split (eager launches): %.2f ms, %.1f M env-steps/s" % (a.elapsed_time(b), c["env_steps"] / a.elapsed_time(b) / 1e3))
# the two kernels on their own
a, b = ev(), ev(); a.record()
for t in range(32): acts, logp = eng.policy(net, counter=5000 + t)
b.record(); torch.cuda.synchronize(); print("policy alone: %.1f us per call" % (a.elapsed_time(b) / 32 * 1e3))
a, b = ev(), ev(); a.record()
for t in range(32): eng.step(acts, reward=rew, done=done)
b.record(); torch.cuda.synchronize(); print("step alone (same actions): %.1f us per call" % (a.elapsed_time(b) / 32 * 1e3))
