"""Host-side rollout logic (deepgroebner_b200/rollout.py) against the reference's own known answers
(deepgroebner/pg.py docstrings) and a direct restatement of TrajectoryBuffer on single trajectories."""
import numpy as np
import pytest
import torch

from deepgroebner_b200.rollout import PairsPolicy, TrajectoryBatch, compute_advantages, discount_rewards


def ref_discount(rewards, gam):
    """pg.discount_rewards (pg.py:18-39) restated for one trajectory."""
    rewards = np.array(rewards, dtype=float)  # the reference computes in float64 (pg.py:34, 74-75)
    out, run = np.zeros(len(rewards)), 0.0
    for i in reversed(range(len(rewards))):
        run = rewards[i] + gam * run
        out[i] = run
    return out


def ref_advantages(rewards, values, gam, lam):
    """pg.compute_advantages (pg.py:42-78)."""
    r, v = np.array(rewards, float), np.array(values, float)
    delta = r - v
    delta[:-1] += gam * v[1:]
    return ref_discount(delta, gam * lam)


def test_docstring_known_answers():
    # pg.py:30-33 and pg.py:68-72
    d = torch.tensor([[False] * 4 + [True]])
    assert discount_rewards(torch.ones(1, 5), d, 0.5)[0].tolist() == [1.9375, 1.875, 1.75, 1.5, 1.0]
    adv = compute_advantages(torch.ones(1, 5), torch.zeros(1, 5), d, 0.5, 0.5)[0]
    assert adv.tolist() == [1.33203125, 1.328125, 1.3125, 1.25, 1.0]


def test_segments_restart_after_done():
    rng = np.random.default_rng(0)
    N, T = 7, 40
    rewards = -rng.integers(1, 9, (N, T)).astype(np.float32)
    values = rng.normal(size=(N, T)).astype(np.float32)
    done = rng.random((N, T)) < 0.15
    rtg = discount_rewards(torch.tensor(rewards), torch.tensor(done), 0.99).numpy()
    adv = compute_advantages(torch.tensor(rewards), torch.tensor(values), torch.tensor(done), 0.99, 0.97).numpy()
    for n in range(N):
        start = 0
        ends = list(np.nonzero(done[n])[0]) + ([T - 1] if not done[n, T - 1] else [])
        for end in ends:
            seg = slice(start, end + 1)
            if done[n, end]:  # complete episode: equals the reference on that trajectory alone
                assert np.allclose(rtg[n, seg], ref_discount(rewards[n, seg], 0.99), rtol=1e-12)
                assert np.allclose(adv[n, seg], ref_advantages(rewards[n, seg], values[n, seg], 0.99, 0.97), rtol=1e-12)
            start = end + 1


def test_trajectory_batch_keeps_only_finished_multi_row_steps():
    N, T = 3, 6
    out = dict(actions=torch.zeros((N, T), dtype=torch.int32), logp=torch.zeros((N, T)),
               reward=-torch.ones((N, T)), lengths=torch.full((N, T), 2, dtype=torch.int32),
               done=torch.zeros((N, T), dtype=torch.uint8),
               obs=torch.arange(N * T * 2 * 2, dtype=torch.int32).reshape(N, T, 2, 2))
    out["done"][0, 2] = 1            # env 0: episode of 3 steps, then an unfinished tail
    out["done"][1, 5] = 1            # env 1: one episode filling the window
    out["lengths"][1, 3] = 1         # a single-row state: filtered like pg.py:201-202
    tb = TrajectoryBatch(out)
    assert tb.finished.tolist() == [[True] * 3 + [False] * 3, [True] * 6, [False] * 6]
    returns, lengths = tb.episode_stats()
    assert sorted(zip(returns.tolist(), lengths.tolist())) == [(-6.0, 6), (-3.0, 3)]
    obs, actions, logp, adv, rtg = tb.get(gam=1.0, lam=1.0, normalize_advantages=False)
    assert obs.shape == (8, 2, 2) and actions.shape == (8,)
    assert sorted(rtg.tolist()) == sorted([-3.0, -2.0, -1.0] + [-6.0, -5.0, -4.0, -2.0, -1.0])
    assert tb.dropped_wide == 0
    # a state with more rows than the stored matrices hold (rollout(pmax=...)) cannot be a training sample
    out["lengths"][0, 1] = 4
    tb = TrajectoryBatch(out)
    obs, actions, logp, adv, rtg = tb.get(gam=1.0, lam=1.0, normalize_advantages=False)
    assert obs.shape == (7, 2, 2) and tb.dropped_wide == 1
    assert sorted(rtg.tolist()) == sorted([-3.0, -1.0] + [-6.0, -5.0, -4.0, -2.0, -1.0])


def test_pairs_policy_reference_forward_masks_padding():
    net = PairsPolicy(cols=4, hidden=32, torch_seed=1)
    obs = torch.tensor([[[1, 2, 0, 3], [0, 1, 1, 0], [-1, -1, -1, -1]]], dtype=torch.int32)
    lp = net.log_probs(obs)
    assert lp.shape == (1, 3) and lp[0, 2] < -1e8
    assert torch.allclose(lp[0, :2].exp().sum(), torch.tensor(1.0), atol=1e-6)
    assert torch.allclose(lp[0, :2], net.log_probs(obs[:, :2])[0], atol=1e-6)
