"""Host logic of the multi-GPU path: block partition of episode ids and the end-of-rollout all-gather of
episode records, exercised with world_size 2 and 3 over gloo on CPU.  The engine is replaced by a stub that fabricates
a record per seed, so the test checks exactly what sharding.py adds: which rank runs which episode and that the
gathered array is in episode order and independent of the number of ranks."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deepgroebner_b200 import _lib, sharding  # noqa: E402

DT = np.dtype(_lib.STATS_DTYPE)


def fake_records(seeds):
    r = np.zeros(len(seeds), DT)
    s = np.asarray(seeds, np.int64)
    r["steps"] = 10 + s % 97
    r["additions"] = 3 * r["steps"] + s % 5
    r["zero_reductions"] = s % 7
    r["nonzero_reductions"] = r["steps"] - r["zero_reductions"]
    r["status"] = 2
    r["trace_hash"] = (s.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    r["discounted_return"] = -0.5 * s
    return r


class StubEngine:
    """run_episodes with the signature sharding.run_sharded uses; records which seeds it was asked to run."""

    def __init__(self):
        self.ran = []
        self.offsets = []

    def set_episode_offset(self, offset=0):
        self.offsets.append(offset)

    def run_episodes(self, strategy, episodes, seed_base, to_host, **kw):
        assert to_host is False
        seeds = np.arange(seed_base, seed_base + episodes)
        self.ran.extend(seeds.tolist())
        rec = fake_records(seeds)
        return torch.from_numpy(rec.view(np.uint8).reshape(-1).copy()), None


def test_shard_bounds_partition():
    for total in (0, 1, 7, 16384, 65536, 65537):
        for ws in (1, 2, 3, 4, 8):
            b = sharding.shard_bounds(total, ws)
            assert b[0] == 0 and b[-1] == total and len(b) == ws + 1
            sizes = np.diff(b)
            assert sizes.max() - sizes.min() <= 1 and (sizes >= 0).all()
            seen = np.concatenate([sharding.shard_seeds(100, total, r, ws) for r in range(ws)])
            assert np.array_equal(seen, np.arange(100, 100 + total))
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_single_process_is_the_identity():
    eng = StubEngine()
    rec = sharding.run_sharded(eng, "degree", 33, seed_base=5)
    assert np.array_equal(rec, fake_records(range(5, 38)))
    s = sharding.summarize(rec)
    assert s["episodes"] == 33 and s["env_steps"] == int(rec["steps"].sum()) and s["finished"] == 33


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        assert sharding.world() == (rank, ws, rank)
        eng = StubEngine()
        rec = sharding.run_sharded(eng, "degree", total, seed_base=1000)
        first, count = sharding.shard_range(total, rank, ws)
        ok = eng.ran == list(range(1000 + first, 1000 + first + count))
        ok = ok and np.array_equal(rec, fake_records(range(1000, 1000 + total)))
        local = sharding.run_sharded(StubEngine(), "degree", total, seed_base=1000, gather=False)
        ok = ok and np.array_equal(local, fake_records(range(1000 + first, 1000 + first + count)))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ws,total", [(2, 64), (2, 65), (3, 100)])
def test_gather_over_gloo(ws, total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, total, q)) for r in range(ws)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(ws))
    assert got == [(r, True) for r in range(ws)]
