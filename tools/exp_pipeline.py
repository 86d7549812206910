"""Diagnostic: a chain of batches with overlapping runners (bb_run on alternating streams, preparation two batches ahead)
against the same batches one after the other; with and without the L2 flush write in front of every runner."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepgroebner_b200.buchberger import BuchbergerEngine, resident_envs

E, K = 16384, 12
eng = BuchbergerEngine("3-20-10-weighted", num_envs=resident_envs(0, 3))
main, alt, side = torch.cuda.current_stream(), torch.cuda.Stream(), torch.cuda.Stream()
seeds = torch.arange(E, dtype=torch.int32, device="cuda")
flush = torch.empty(int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 256 << 20, dtype=torch.uint8, device="cuda")
outs = [torch.empty(E * 72, dtype=torch.uint8, device="cuda") for _ in range(2)]
ev = lambda: torch.cuda.Event(enable_timing=True)

def prep():
    with torch.cuda.stream(side):
        eng.prepare_episodes(E, seeds=seeds)

def chain(two_streams, do_flush, ahead):
    torch.cuda.synchronize()
    for _ in range(ahead):
        prep()
    torch.cuda.synchronize()
    t0, t1 = ev(), ev()
    t0.record(main); alt.wait_event(t0); side.wait_event(t0)
    marks = []
    for i in range(K):
        with torch.cuda.stream([main, alt][i % 2] if two_streams else main):
            if do_flush:
                flush.fill_(1)
            a, b = ev(), ev(); a.record()
            eng.run_episodes("degree", episodes=E, seeds=seeds, to_host=False, out=outs[i % 2])
            b.record(); marks.append((a, b))
        if ahead:
            prep()
    main.wait_stream(alt); main.wait_stream(side); t1.record(main)
    torch.cuda.synchronize()
    span = t0.elapsed_time(t1)
    for _ in range(ahead):
        eng.run_episodes("degree", episodes=E, seeds=seeds, to_host=False)
    torch.cuda.synchronize()
    return round(span / K, 4), [(round(t0.elapsed_time(a), 2), round(t0.elapsed_time(b), 2)) for a, b in marks[:6]]

for name, cfg in (("one stream, prepare inline, no flush", (False, False, 0)), ("one stream, prepare ahead, no flush", (False, False, 2)),
                  ("two streams, prepare ahead, no flush", (True, False, 2)), ("two streams, prepare ahead, flush", (True, True, 2)),
                  ("one stream, prepare ahead, flush", (False, True, 2))):
    chain(*cfg)
    print(name, chain(*cfg))

# ---- the end-to-end loop of bench.py with device-side marks: where does a step of the host-paced pipeline go?
import time
import numpy as np
from deepgroebner_b200 import _lib
seeds_host = torch.arange(E, dtype=torch.int32).pin_memory()
host_out = [torch.empty(E * 72, dtype=torch.uint8).pin_memory() for _ in range(2)]
def prep_host():
    with torch.cuda.stream(side):
        return eng.prepare_episodes(E, seeds=seeds_host)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = ev(); t0.record(main); alt.wait_event(t0); side.wait_event(t0)
    w0 = time.perf_counter()
    ahead = [prep_host(), prep_host()]
    done, marks, host_t = [], [], []
    for i in range(K):
        with torch.cuda.stream([main, alt][i % 2]):
            a, b, c = ev(), ev(), ev(); a.record()
            buf, _ = eng.run_episodes("degree", episodes=E, seeds=ahead[i], to_host=False, out=outs[i % 2])
            b.record()
            host_out[i % 2].copy_(buf, non_blocking=True)
            c.record(); done.append(c); marks.append((a, b, c))
        ahead.append(prep_host())
        host_t.append(round((time.perf_counter() - w0) * 1e3, 2))
        if i >= 1:
            done[i - 1].synchronize()
            host_out[(i - 1) % 2].numpy().view(np.dtype(_lib.STATS_DTYPE))["steps"][:E].sum()
    done[-1].synchronize()
    wall = (time.perf_counter() - w0) * 1e3
    torch.cuda.synchronize()
    print("e2e loop: %.4f ms/step; host enqueue times %r; device (start, kernel end, copy end) %r" % (
        wall / K, host_t[:6], [(round(t0.elapsed_time(a), 2), round(t0.elapsed_time(b), 2), round(t0.elapsed_time(c), 2)) for a, b, c in marks[:6]]))
    for _ in range(2):
        eng.run_episodes("degree", episodes=E, seeds=seeds, to_host=False)
