"""Diagnostic: does the preparation of batch i + 1 (side stream) co-run with the runner of batch i?  Times each stream's
work with its own CUDA events, alone and together."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepgroebner_b200.buchberger import BuchbergerEngine, resident_envs

E = 16384
SLOTS = int(sys.argv[1]) if len(sys.argv) > 1 else resident_envs(0, 3)
print("slots", SLOTS)
eng = BuchbergerEngine("3-20-10-weighted", num_envs=SLOTS)
main, side = torch.cuda.current_stream(), torch.cuda.Stream()
seeds = torch.arange(E, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)

def run():
    eng.run_episodes("degree", episodes=E, seeds=seeds, to_host=False)

for _ in range(3):
    run()
torch.cuda.synchronize()
for mode in ("run alone (prepared)", "together"):
    res = []
    for rep in range(4):
        with torch.cuda.stream(side):
            eng.prepare_episodes(E, seeds=seeds)      # the batch the runner below consumes
        torch.cuda.synchronize()
        flush.fill_(1)
        a, b, c, d = ev(), ev(), ev(), ev()
        a.record(main)
        side.wait_event(a)
        if mode != "run alone (prepared)":
            with torch.cuda.stream(side):
                if mode.endswith("late"):
                    torch.cuda._sleep(400000)
                c.record(side)
                eng.prepare_episodes(E, seeds=seeds)   # the NEXT batch (other staging set)
                d.record(side)
        if mode != "prepare alone":
            run()
        b.record(main)
        torch.cuda.synchronize()
        res.append((round(a.elapsed_time(b), 3), round(c.elapsed_time(d), 3) if mode != "run alone (prepared)" else None,
                    round(a.elapsed_time(d), 3) if mode != "run alone (prepared)" else None))
        if mode == "prepare alone":
            run()   # consume, keeps the staging sets alternating
            torch.cuda.synchronize()
    print(mode, "(main a->b, side c->d, a->d):", res)
