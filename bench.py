#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched Buchberger environment on B200 (BASELINE.json metric), beside the
reference's own CPU environment.

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]      the unmodified reference on the host cores

Workload (BASELINE.json configs[1]): 3-20-10-weighted, 16384 episodes per GPU run to completion under Degree
selection; episode e draws its ideal from the reference generator stream seed(e).  One bench "step" = one pass of
the hot path over that batch = one launch of the persistent episode kernel (reset + select + spoly + reduce +
update for every step of every episode).  `value` = env steps / device time with the inputs (seeds) resident in
HBM; `e2e` = the same through BuchbergerEngine.run_episodes with HOST buffers (pinned seeds H2D, episode records
D2H, inside the timed region).  The oracle / reference is only ever used here as the cpu_baseline leg, the
`--impl reference` arm and a post-run spot check -- never inside a timed GPU region.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
# version banner there under torchrun), so fd 1 is pointed at stderr for the whole run and the JSON line goes to a
# private duplicate of the original stdout.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
sys.stdout = sys.stderr


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


DIST = "3-20-10-weighted"
STRATEGY = "degree"
EPISODES = 16384
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"

# name -> (ideal_dist, selection, episodes, "weak": per GPU | "strong": in total, BASELINE.json config it is)
WORKLOADS = {
    "episodes": ("3-20-10-weighted", "degree", 16384, "weak", "configs[1]"),
    "u3": ("3-20-10-uniform", "degree", 65536, "strong", "configs[2]"),
    "u5": ("5-5-10-uniform", "degree", 65536, "strong", "configs[2]"),
    "cyclic6": ("cyclic-6", "random", 1024, "weak", "configs[4]"),
}
SCALING, CONFIG_ID, DATA_NOTE = "weak", "configs[1]", ""
SEL_SEED = 1234  # Random selection: episode e draws choice() from minstd_rand0 seeded SEL_SEED + e


def algorithmic_bytes(c):
    """12 B per term read/written by an addition, 8 B per reducer lead monomial examined,
    24 B per lead term moved to the remainder, 8 B per basis / pair entry touched by update()."""
    return (12 * (c["terms_read"] + c["terms_written"]) + 8 * c["lms_scanned"] + 24 * c["term_moves"]
            + 8 * (c["update_basis"] + c["update_pairs"]))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_capture():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return {}
    return {}


def ncu_traffic():
    """DRAM bytes per launch of k_run from the committed ncu --set full capture, if any."""
    return ncu_capture().get("k_run_dram_bytes_per_launch")


def int_pipe(launch_s, sm_mhz, sm_count):
    """The secondary roofline BASELINE.json allows for this path (instruction issue / integer pipe):
    warp instructions per launch of k_run (a property of the kernel + workload, from the committed ncu capture of the
    same command) over the LIVE launch time, against the issue peak = SMs x 4 schedulers x SM clock sampled during
    the timed region.  The pipe percentages are the capture's."""
    c = ncu_capture()
    inst = c.get("k_run_warp_instructions_per_launch")
    if not inst or not sm_mhz:
        return None
    peak = sm_count * 4 * sm_mhz * 1e6
    ach = inst / launch_s
    return {"bound": "issue", "achieved": ach / 1e9, "peak": peak / 1e9, "unit": "G warp-inst/s", "frac": ach / peak,
            "warp_instructions_per_launch": inst, "sm_mhz": sm_mhz,
            "alu_pipe_pct_of_peak_while_active": c.get("k_run_alu_pipe_pct_active"),
            "issue_slots_busy_pct": c.get("k_run_issue_active_pct"),
            "threads_per_warp_instruction": c.get("k_run_threads_per_warp_instruction"),
            "source": c.get("source")}


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every ~2 ms through NVML (in-process thread) while the timed
    region runs -- the region is tens of milliseconds, far below nvidia-smi's sampling period."""

    def __init__(self, torch_index):
        import threading
        self.samples, self.reasons, self.max_mhz, self.err = [], set(), None, None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(torch_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_index)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        except Exception as ex:  # no NVML: report that instead of inventing clocks
            self.err = repr(ex)

    def _loop(self):
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                 ("hw_power_brake", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown))
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for n, bit in names:
                    if r & bit:
                        self.reasons.add(n)
            except Exception as ex:
                self.err = repr(ex)
                return
            time.sleep(0.002)

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.samples:
            sm = sorted(self.samples)
            out["sm_mhz"] = sm[len(sm) // 2]
        if self.err:
            out["error"] = self.err
        return out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def load_cpu_oracle():
    from oracle import oracle as O
    if O.have_ref():
        return O.load_ref(), "reference"
    return O.load_port(), "port"


def cpu_sample(orc, kind, seconds, threads):
    """Times the reference env (reset + Degree select + step to completion) on `threads` host threads over a
    bounded sample of the SAME workload (episodes seed 0..count-1), sized for about `seconds` of wall time."""
    if kind != "reference":
        threads = 1  # the C restatement is single-threaded
    n0 = (64 if STRATEGY != "random" else 1) * threads
    probe = cpu_run(orc, kind, 0, n0, threads)
    rate = probe["steps"] / max(probe["seconds"], 1e-9)
    steps_per_ep = probe["steps"] / float(n0)
    count = int(min(EPISODES, max(n0, seconds * rate / steps_per_ep)))
    if count > EPISODES // 2:
        count = EPISODES  # the whole workload fits the budget: no sampling at all
    r = cpu_run(orc, kind, 0, count, threads)
    if count == EPISODES and r["seconds"] < seconds / 2:
        # the whole workload is quicker than the time budget: repeat it so the figure is not a 0.2 s measurement
        reps = int(min(64, max(1, seconds / max(r["seconds"], 1e-3)))) - 1
        for _ in range(reps):
            q = cpu_run(orc, kind, 0, count, threads)
            for k in ("steps", "additions", "seconds"):
                r[k] += q[k]
        r["reps"] = reps + 1
    return r, count, threads


def cpu_run(orc, kind, seed0, count, threads):
    """Episodes seed0 .. seed0+count-1 of the current workload on the host: the reference env stepped with the
    selection comparators (First/Degree/Normal/Sugar), or the reference's own buchberger() loop for seeded Random."""
    if STRATEGY == "random":
        if kind != "reference":
            raise SystemExit("the Random-selection CPU arm needs oracle/_ref (the unmodified reference)")
        return orc.bench_buchberger(DIST, STRATEGY, seed0, count, nthreads=threads, sel_seed0=SEL_SEED + seed0)
    return orc.bench_selection(DIST, STRATEGY, seed0, count, nthreads=threads)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc, kind = load_cpu_oracle()
    threads = host_threads()
    # each step = the first `count` episodes of the workload, sized for ~4 s per step
    probe, count, threads = cpu_sample(orc, kind, 4.0, threads)
    for _ in range(args.warmup):
        cpu_run(orc, kind, 0, max(count // 8, threads), threads)
    steps = adds = 0
    secs = 0.0
    for _ in range(args.steps):
        r = cpu_run(orc, kind, 0, count, threads)
        steps += r["steps"]; adds += r["additions"]; secs += r["seconds"]
    value = steps / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True,
        "scaling": SCALING, "vs_baseline": None, "dtype": "int32 (GF(32003) coefficients, int exponent vectors)",
        "data": DATA_NOTE.replace("on-device restatement of the reference", "reference"),
        "config": {"workload": "%s, %s selection, episodes to completion; bounded sample: episodes 0..%d of %d per step"
                               % (DIST, STRATEGY, count - 1, EPISODES), "episodes_per_step": count,
                   "host_threads": threads},
        "additions_per_sec": adds / secs,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "%d episodes (seeds 0..%d) x %d steps, one BuchbergerEnv per thread, "
                                   "env.seed(e); reset(); %s until P empty"
                                   % (count, count - 1, args.steps,
                                      "the reference buchberger() loop with seeded Random selection" if STRATEGY == "random"
                                      else STRATEGY + " select + step")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from deepgroebner_b200 import _lib
    from deepgroebner_b200.buchberger import BuchbergerEngine, resident_envs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    from deepgroebner_b200 import sharding
    from deepgroebner_b200.ideals import FixedIdealGenerator, parse_ideal_dist
    spec = parse_ideal_dist(DIST, 32003)
    nvars = spec.nvars() if isinstance(spec, FixedIdealGenerator) else spec.n
    if SCALING == "weak":   # every rank runs its own EPISODES episodes, disjoint seeds
        ep_first, ep_local = rank * EPISODES, EPISODES
    else:                   # EPISODES in total, contiguous blocks (deepgroebner_b200/sharding.py)
        ep_first, ep_local = sharding.shard_range(EPISODES, rank, world)
    slots = args.slots or min(resident_envs(local, nvars), ep_local)
    eng = BuchbergerEngine(DIST, num_envs=slots, device="cuda:%d" % local)
    seeds_host = torch.arange(ep_first, ep_first + ep_local, dtype=torch.int32).pin_memory()
    seeds_dev = seeds_host.to(dev)
    stats_bytes = ep_local * 72
    stats_dev = torch.empty(stats_bytes, dtype=torch.uint8, device=dev)
    stats_host = torch.empty(stats_bytes, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    import ctypes as C
    lib = eng.lib

    def launch(gb):
        rc = lib.bb_run(eng.h, _lib.SELECTION[STRATEGY], ep_local, 0, C.c_void_p(seeds_dev.data_ptr()), SEL_SEED + ep_first,
                        0, 0.99, gb, C.c_void_p(stats_dev.data_ptr()), None, 0, 0,
                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc < 0:
            raise RuntimeError(lib.bb_last_error(eng.h))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        launch(0)
    torch.cuda.synchronize()
    eng.counters(reset=True)

    # ---- device-timed region: K launches, L2 flushed between them (flush outside the event pairs)
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(1)
        a.record()
        launch(0)
        b.record()
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    counters = eng.counters(reset=True)
    stats = stats_dev.cpu().numpy().view(np.dtype(_lib.STATS_DTYPE))
    assert (stats["status"] == 2).all(), "not every episode finished: status histogram %r" % (
        dict(zip(*[x.tolist() for x in np.unique(stats["status"], return_counts=True)])),)
    steps_per_launch = int(stats["steps"].sum())
    adds_per_launch = int(stats["additions"].sum())
    assert counters["env_steps"] == steps_per_launch * args.steps

    # ---- end-to-end through the public API with host buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        seeds_dev.copy_(seeds_host, non_blocking=True)
        launch(0)
        stats_host.copy_(stats_dev, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([dev_ms, e2e_s * 1000.0], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(steps_per_launch), float(adds_per_launch)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max = float(t[0]), float(t[1])
    total_steps, total_adds = float(tot[0]), float(tot[1])
    per_rank = None
    if world > 1:   # diagnosis of the scaling figure: every rank's own device time, fastest single launch and work
        mine = torch.tensor([dev_ms / args.steps, min(a.elapsed_time(b) for a, b in ev), float(steps_per_launch)],
                            dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [round(float(x[0]), 4) for x in allr], "min_ms": [round(float(x[1]), 4) for x in allr],
                    "env_steps_per_launch": [int(x[2]) for x in allr]}

    if rank == 0:
        # spot check (outside every timed region): a sample of the timed output against the CPU oracle
        orc, kind = load_cpu_oracle()
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from hashing import trace_hash
        env = orc.env(DIST)
        for e in range(0, ep_local, max(1, ep_local // (4 if STRATEGY == "random" else 16))):
            env.seed(ep_first + e)
            F, _ = env.reset()
            if STRATEGY == "random":   # the reference's own buchberger() loop with the episode's selection seed
                _, st = orc.buchberger(F, selection="random", gamma=0.99, seed=SEL_SEED + ep_first + e)
                ok = (stats["steps"][e] == st["zero_reductions"] + st["nonzero_reductions"]
                      and stats["additions"][e] == st["polynomial_additions"]
                      and stats["discounted_return"][e] == st["discounted_return"])
            else:
                tr = env.run(selection=STRATEGY)
                ok = stats["steps"][e] == len(tr) and int(stats["trace_hash"][e]) == trace_hash(tr)
            assert ok, "GPU episode %d differs from the %s oracle" % (ep_first + e, kind)

        peak, peak_src = measured_peak()
        abytes = algorithmic_bytes(counters) / args.steps
        launch_s = dev_ms / 1000.0 / args.steps
        achieved = abytes / launch_s / 1e9
        value = total_steps * args.steps / (dev_ms_max / 1000.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": SCALING, "vs_baseline": None,
            "dtype": "u64 packed monomials + u32 GF(32003) coefficients (integer)",
            "data": DATA_NOTE,
            "config": {"workload": "%s, %d episodes %s to completion, %s selection (BASELINE %s)"
                                   % (DIST, EPISODES, "per GPU" if SCALING == "weak" else "in total, sharded", STRATEGY,
                                      CONFIG_ID),
                       "episodes_per_gpu": ep_local, "env_steps_per_launch": steps_per_launch, "slots": slots,
                       "parallelism": "episodes sharded across GPUs, no collective on the step path",
                       "l2": "256 MiB flush write between timed launches"},
            "additions_per_sec": total_adds * args.steps / (dev_ms_max / 1000.0),
            "spair_reductions_per_sec": value,
            # per step: k_prepare + k_order + k_run (bb_run, one batch) -- the L2 flush fill and the queue memset are not ours
            "gpu_launches": 3 * args.steps * ((ep_local + 65535) // 65536),
            "wall_s_timed_region": wall,
            "clocks": clocks,
            "e2e": {"value": total_steps * args.steps / (e2e_ms_max / 1000.0), "unit": UNIT,
                    "h2d_bytes_per_step": ep_local * 4, "d2h_bytes_per_step": stats_bytes},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "peak_source": peak_src, "kernel": "k_run",
                         "algorithmic_bytes_per_launch": abytes,
                         "note": "latency/integer bound at binomial sizes (working set lives in L1/L2); see DESIGN.md"},
            "counters_per_launch": {k: v / args.steps for k, v in counters.items()},
        }
        if per_rank:
            line["per_rank"] = per_rank
        ip = int_pipe(launch_s, (clocks or {}).get("sm_mhz"), eng.sm_count) if args.workload == "episodes" else None
        if ip:
            line["int_pipe"] = ip
        if world == 1 and not args.no_cpu:
            threads = host_threads()
            r, count, threads = cpu_sample(orc, kind, args.cpu_seconds, threads)
            line["cpu_baseline"] = {
                "value": r["steps"] / r["seconds"], "unit": UNIT, "cores": threads, "kind": kind,
                "sample": "%d episodes (seeds 0..%d) of the same workload x %d passes, one reference BuchbergerEnv per "
                          "host thread, %.1f s" % (count, count - 1, r.get("reps", 1), r["seconds"])}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def rollout_arm(args):
    """BASELINE configs[3]: LeadMonomialsEnv(k=2) PPO rollout -- env step + PMLP(128) policy head sampling on device,
    fused in one launch (bb_rollout), auto-reset.  Extra workload (`--workload rollout`); the default stays configs[1]."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from deepgroebner_b200 import LeadMonomialsEnv
    from deepgroebner_b200.rollout import PairsPolicy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    N, T = args.envs, args.horizon
    env = LeadMonomialsEnv(DIST, k=2, num_envs=N, device="cuda:%d" % local, pmax=64)
    env.seed(np.arange(rank * N, rank * N + N))
    env.engine.set_auto_reset(True)
    env.engine.reset()
    net = PairsPolicy(env.engine.cols, 128, torch_seed=0, seed=1, device=dev)
    out = env.engine.rollout(net, T)
    host = {k: torch.empty_like(out[k], device="cpu").pin_memory() for k in ("reward", "done")}
    w_host = [t.cpu().pin_memory() for t in net.parameters()]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for i in range(max(args.warmup, 3)):
        flush.fill_(1)
        env.engine.rollout(net, T, counter0=i * T, out=out)
    torch.cuda.synchronize()
    env.engine.counters(reset=True)
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i, (a, b) in enumerate(ev):
        flush.fill_(1)
        a.record()
        env.engine.rollout(net, T, counter0=(100 + i) * T, out=out)
        b.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    counters = env.engine.counters(reset=True)
    steps_done = counters["env_steps"]
    # end to end: weights from pinned host memory, rewards + done flags back to the host
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        for t, h in zip(net.parameters(), w_host):
            t.copy_(h, non_blocking=True)
        env.engine.rollout(net, T, counter0=(200 + i) * T, out=out)
        for k in host:
            host[k].copy_(out[k], non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_steps = env.engine.counters(reset=True)["env_steps"]
    t = torch.tensor([dev_ms, e2e_s * 1000.0], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(steps_done), float(e2e_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        peak, peak_src = measured_peak()
        abytes = (algorithmic_bytes(counters) + 4 * 12 * counters["obs_rows"]) / args.steps
        achieved = abytes / (dev_ms / 1000.0 / args.steps) / 1e9
        line = {
            "metric": "rollout_env_steps_per_sec", "value": float(tot[0]) / (float(t[0]) / 1000.0), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": float(t[0]) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 packed monomials + u32 GF(32003) coefficients (integer); fp32 policy head",
            "data": "synthetic (device generator, env e = stream seed e; PMLP(128) Glorot-initialised, torch seed 0)",
            "config": {"workload": "%s LeadMonomialsEnv(k=2), %d envs x %d fused steps per launch, PMLP(128) sampling on "
                                   "device, auto-reset (BASELINE configs[3])" % (DIST, N, T),
                       "envs_per_gpu": N, "horizon": T, "l2": "256 MiB flush write between timed launches"},
            "gpu_launches": args.steps, "clocks": clocks,   # one fused k_rollout per step
            "e2e": {"value": float(tot[1]) / (float(t[1]) / 1000.0), "unit": UNIT,
                    "h2d_bytes_per_step": sum(x.numel() * 4 for x in w_host),
                    "d2h_bytes_per_step": sum(h.numel() * h.element_size() for h in host.values())},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "k_rollout",
                         "algorithmic_bytes_per_launch": abytes},
        }
        if world == 1 and not args.no_cpu:
            orc, kind = load_cpu_oracle()
            if kind == "reference":
                threads = host_threads()
                r = orc.bench_random(DIST, 0, 2000 * threads, nthreads=threads)
                line["cpu_baseline"] = {"value": r["steps"] / r["seconds"], "unit": UNIT, "cores": threads, "kind": kind,
                                        "sample": "scripts/random_episodes.cpp loop (uniform-random actions, no network) on "
                                                  "LeadMonomialsEnv(k=2), %d episodes, %.1f s" % (2000 * threads, r["seconds"])}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="episodes", choices=sorted(WORKLOADS) + ["rollout"],
                    help="episodes = BASELINE configs[1] (the headline); u3/u5 = configs[2]; cyclic6 = configs[4]; "
                         "rollout = configs[3]")
    ap.add_argument("--episodes", type=int, default=0, help="override the workload's episode count")
    ap.add_argument("--envs", type=int, default=16384, help="rollout workload: environments per GPU")
    ap.add_argument("--horizon", type=int, default=128, help="rollout workload: fused steps per launch")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--slots", type=int, default=0, help="environment slots (0 = one resident wave)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    global DIST, STRATEGY, EPISODES, SCALING, CONFIG_ID, DATA_NOTE
    DIST, STRATEGY, EPISODES, SCALING, CONFIG_ID = WORKLOADS.get(args.workload, WORKLOADS["episodes"])
    if args.episodes:
        EPISODES = args.episodes
    DATA_NOTE = ("synthetic (fixed ideal staged on device; episode e = Random-selection stream seed %d + e)" % SEL_SEED
                 if STRATEGY == "random" else
                 "synthetic (on-device restatement of the reference RandomBinomialIdealGenerator; episode e = stream seed e)")
    if args.impl == "reference":
        reference_arm(args)
    elif args.workload == "rollout":
        rollout_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
