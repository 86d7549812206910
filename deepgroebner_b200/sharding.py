"""Episodes across the GPUs of one box.

Episodes are independent -- the reference environment has no shared state (buchberger.h:161-208) -- so the path
shards with NO collective on the step path: one process per GPU, rank r owns a contiguous block of episode ids
(= ideal-stream seeds) and runs it on its own handle.  The only exchange is the optional end-of-rollout
all-gather of the per-episode records (bb_episode_stats: BuchbergerStats of buchberger.h:99-105 plus checksums),
72 bytes per episode; NCCL over NVLink on the GPU box, gloo in the CPU tests.

Nothing here touches the device except through the tensors the caller hands in.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

RECORD_BYTES = np.dtype(_lib.STATS_DTYPE).itemsize


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(total, world_size):
    """Start offsets [world_size + 1] of the contiguous blocks: the first total % world_size ranks get one more."""
    if total < 0 or world_size < 1:
        raise ValueError("need total >= 0 and world_size >= 1")
    q, r = divmod(total, world_size)
    sizes = [q + (1 if i < r else 0) for i in range(world_size)]
    return [0] + list(np.cumsum(sizes).astype(int))


def shard_range(total, rank, world_size):
    """(first episode id, count) of rank's block."""
    if not 0 <= rank < world_size:
        raise ValueError("rank outside [0, world_size)")
    b = shard_bounds(total, world_size)
    return int(b[rank]), int(b[rank + 1] - b[rank])


def shard_seeds(seed_base, total, rank, world_size):
    """The int32 ideal-stream seeds of rank's block when episode e uses seed seed_base + e."""
    first, count = shard_range(total, rank, world_size)
    return np.arange(seed_base + first, seed_base + first + count, dtype=np.int32)


def gather_records(local, total, group=None):
    """All-gather of the per-episode records: `local` is this rank's uint8 tensor [count * 72] (or a structured numpy
    array) for its block of shard_range(total, rank, world); returns the numpy structured array [total] in episode
    order on EVERY rank.  Blocks may differ in size by one: they are padded to the largest for the collective."""
    if isinstance(local, np.ndarray):
        local = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).reshape(-1))
    local = local.contiguous().view(-1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = local.cpu().numpy()
        assert out.size == total * RECORD_BYTES
        return out.view(np.dtype(_lib.STATS_DTYPE))
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(total, ws)
    first, count = shard_range(total, rank, ws)
    assert local.numel() == count * RECORD_BYTES, "local block has %d bytes, expected %d" % (local.numel(), count * RECORD_BYTES)
    width = max(bounds[i + 1] - bounds[i] for i in range(ws)) * RECORD_BYTES
    send = torch.zeros(width, dtype=torch.uint8, device=local.device)
    send[:local.numel()] = local
    recv = torch.empty(ws * width, dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(ws, width)
    parts = [recv[i, :(bounds[i + 1] - bounds[i]) * RECORD_BYTES] for i in range(ws)]
    return np.concatenate(parts).view(np.dtype(_lib.STATS_DTYPE))


def run_sharded(engine, strategy, total, seed_base=0, gather=True, group=None, **run_kwargs):
    """Runs episodes seed_base .. seed_base + total - 1 over all ranks: this rank runs its block on `engine`
    (BuchbergerEngine.run_episodes -> bb_run).  Returns the structured record array of all `total` episodes when
    gather (identical on every rank), else of this rank's block only.  Episode e's record does not depend on the
    number of ranks (tests/test_sharding.py, tests/test_gpu_parity.py)."""
    if dist.is_available() and dist.is_initialized():
        rank, ws = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, ws = 0, 1
    first, count = shard_range(total, rank, ws)
    # the selection seed of 'random' and the staged ideal an episode replays follow the GLOBAL episode index
    engine.set_episode_offset(first)
    try:
        buf, _ = engine.run_episodes(strategy, episodes=count, seed_base=seed_base + first, to_host=False, **run_kwargs)
    finally:
        engine.set_episode_offset(0)
    buf = buf[:count * RECORD_BYTES]
    if not gather:
        return buf.cpu().numpy().view(np.dtype(_lib.STATS_DTYPE))
    return gather_records(buf, total, group)


def summarize(records):
    """Totals the reference's strategy tables report (scripts/make_strat.cpp:51-69): sums over episodes."""
    return {"episodes": int(records.shape[0]), "env_steps": int(records["steps"].sum()),
            "additions": int(records["additions"].sum()), "zero_reductions": int(records["zero_reductions"].sum()),
            "nonzero_reductions": int(records["nonzero_reductions"].sum()),
            "finished": int((records["status"] == 2).sum())}
