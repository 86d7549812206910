"""Diagnosis (timing build, `make -C deepgroebner_b200/csrc variant V=timing VFLAGS=-DBBW_TIMING`): cycles per phase of an
addition of the CTA-per-environment runner, for warp 0 (holds reducer terms) and warp 7 (searches), one cyclic-6 episode."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepgroebner_b200.buchberger import BuchbergerEngine
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1237
eng = BuchbergerEngine("cyclic-6", num_envs=1)
eng.lib.bb_debug_read.argtypes = [C.c_void_p]
buf = np.zeros(64, np.uint64)
stats, _ = eng.run_episodes("random", episodes=1, selection_seed=seed)
eng.lib.bb_debug_read(C.c_void_p(buf.ctypes.data))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); stats, _ = eng.run_episodes("random", episodes=1, selection_seed=seed); b.record(); torch.cuda.synchronize()
eng.lib.bb_debug_read(C.c_void_p(buf.ctypes.data))
adds = int(stats["additions"][0]); ms = a.elapsed_time(b)
names = ["loop/sync staging", "rank+mark B", "predict+search", "barrier 1", "collect+head+clear", "write slots", "stage next", "barrier 2",
         "step bookkeeping", "select+take+spoly stage", "before unpredicted search", "unpredicted search", "loop exit", "update()", "", ""]
print("%d additions, %.1f ms, %.2f us each" % (adds, ms, 1000 * ms / adds))
for w, base in (("warp 0", 0), ("warp 7", 16)):
    tot = sum(int(buf[base + i]) for i in range(14))
    print(w, "cycles per addition: total %.0f" % (tot / adds), " ".join("%s %.0f;" % (names[i], int(buf[base + i]) / adds) for i in range(14)))
