"""GPU tests of the fused policy head and rollout.  The head is floating point: it is
compared with the plain torch fp32 forward of the same network (PairsPolicy.log_probs, the restatement of
networks.py:522-571) at rtol = atol = 1e-5; everything the environment does under the sampled actions stays
bit-exact against the oracle."""
import numpy as np
import pytest

from hashing import hash_item

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-5, atol=1e-5)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def best_oracle():
    from oracle import oracle as O
    return O.load_ref() if O.have_ref() else O.load_port()


def device_uniform(seed, env, counter):
    return float(int(hash_item(np.uint64(seed + env), np.uint64(counter))) >> 40) * 2.0 ** -24


@pytest.mark.parametrize("dist,k,hidden", [("3-20-10-weighted", 2, 128), ("3-20-10-weighted", 1, 32),
                                           ("5-5-10-uniform", 2, 64), ("5-5-10-uniform", 3, 256)])
def test_policy_head_matches_torch_fp32(torch_cuda, dist, k, hidden):
    torch = torch_cuda
    from deepgroebner_b200 import LeadMonomialsEnv
    from deepgroebner_b200.rollout import PairsPolicy
    N, pmax = 128, 256
    env = LeadMonomialsEnv(dist, k=k, num_envs=N, pmax=pmax)
    env.seed(70)
    obs, lengths = env.reset()
    net = PairsPolicy(env.engine.cols, hidden, torch_seed=3, seed=99, device="cuda")
    net.b1 = torch.linspace(-0.3, 0.3, hidden, device="cuda")
    net.b2 = torch.tensor([0.25], device="cuda")
    rng = np.random.default_rng(0)
    for step in range(6):
        actions, logp, allp = env.engine.policy(net, counter=step, return_all=True, pmax=pmax)
        ref = net.log_probs(obs)                        # [N, pmax], padded rows ~ -1e9
        lens = lengths.cpu().numpy()
        a, lp, ap, rf = actions.cpu().numpy(), logp.cpu().numpy(), allp.cpu().numpy(), ref.cpu().numpy()
        for e in range(N):
            n = lens[e]
            if n == 0:
                assert a[e] == 0
                continue
            assert np.allclose(ap[e, :n], rf[e, :n], **TOL), (step, e)
            assert np.isneginf(ap[e, n:]).all()
            assert 0 <= a[e] < n and lp[e] == ap[e, a[e]]
            # the sampled row is the inverse CDF of the reference distribution at the documented uniform
            p = np.exp(rf[e, :n].astype(np.float64))
            cdf = np.cumsum(p) / p.sum()
            u = device_uniform(99, e, step)
            assert cdf[a[e]] > u - 1e-4 and (a[e] == 0 or cdf[a[e] - 1] <= u + 1e-4), (step, e, u, a[e])
        g_actions, g_logp = env.engine.policy(net, counter=step, greedy=True)
        ga = g_actions.cpu().numpy()
        for e in range(N):
            if lens[e]:
                assert rf[e, ga[e]] >= rf[e, :lens[e]].max() - 1e-5
        acts = np.array([rng.integers(l) if l else 0 for l in lens], dtype=np.int32)
        (obs, lengths), _, _, _ = env.step(torch.as_tensor(acts, device="cuda"))


def test_auto_reset_step_continues_the_ideal_stream(torch_cuda):
    """bb_set_auto_reset: done = 1 on the finishing transition, then the first state of the next ideal of the same
    stream -- exactly what the reference loop `if done: state = env.reset()` produces."""
    torch = torch_cuda
    from deepgroebner_b200 import LeadMonomialsEnv
    orc = best_oracle()
    N = 32
    env = LeadMonomialsEnv("3-20-10-weighted", k=2, num_envs=N, pmax=256)
    env.engine.set_auto_reset(True)
    env.seed(np.arange(200, 200 + N))
    refs = []
    for e in range(N):
        r = orc.lm_env("3-20-10-weighted", k=2)
        r.seed(200 + e)
        refs.append(r)
    obs, lengths = env.reset()
    states = [r.reset() for r in refs]
    resets = 0
    for step in range(150):
        o, lens = obs.cpu().numpy(), lengths.cpu().numpy()
        for e in range(N):
            assert lens[e] == len(states[e]) > 0 and np.array_equal(o[e, :lens[e]], states[e])
        acts = np.zeros(N, np.int32)                  # First selection: short episodes, many resets
        (obs, lengths), reward, done, _ = env.step(torch.as_tensor(acts, device="cuda"))
        rw, dn = reward.cpu().numpy(), done.cpu().numpy()
        for e in range(N):
            s, r, d, _ = refs[e].step(0)
            assert rw[e] == r and bool(dn[e]) == d
            if d:
                s = refs[e].reset()
                resets += 1
            states[e] = s
    assert resets > N


def test_fused_rollout_replays_on_the_oracle(torch_cuda):
    """bb_rollout: T fused steps (policy, sample, step, auto-reset).  Replaying the recorded actions through oracle
    environments reproduces every reward, done flag, pair count and state matrix; the recorded log-probabilities
    equal the torch forward on the recorded states."""
    torch = torch_cuda
    from deepgroebner_b200 import LeadMonomialsEnv
    from deepgroebner_b200.rollout import PairsPolicy, collect
    orc = best_oracle()
    N, T, pmax = 48, 160, 128
    env = LeadMonomialsEnv("3-20-10-weighted", k=2, num_envs=N, pmax=pmax)
    env.seed(np.arange(10, 10 + N))
    env.reset()
    net = PairsPolicy(env.engine.cols, 128, torch_seed=11, seed=5, device="cuda")
    tb = collect(env, net, T, counter0=1000, store_obs=True, pmax=pmax)
    A, LP, R = tb.actions.cpu().numpy(), tb.logp.cpu().numpy(), tb.reward.cpu().numpy()
    D, L, O = tb.done.cpu().numpy(), tb.lengths.cpu().numpy(), tb.obs.cpu().numpy()
    ref_lp = net.log_probs(tb.obs.reshape(N * T, pmax, -1)).reshape(N, T, pmax).cpu().numpy()
    episodes = 0
    for e in range(N):
        r = orc.lm_env("3-20-10-weighted", k=2)
        r.seed(10 + e)
        s = r.reset()
        for t in range(T):
            assert L[e, t] == len(s) and np.array_equal(O[e, t, :len(s)], s) and (O[e, t, len(s):] == -1).all()
            a = A[e, t]
            assert 0 <= a < len(s)
            assert abs(LP[e, t] - ref_lp[e, t, a]) <= 1e-5 + 1e-5 * abs(ref_lp[e, t, a])
            s, rew, done, _ = r.step(int(a))
            assert R[e, t] == rew and bool(D[e, t]) == done
            if done:
                s = r.reset()
                episodes += 1
    assert episodes > 0
    rets, lens = tb.episode_stats()
    assert len(rets) == episodes and int(lens.sum()) == int(tb.finished.sum())
    # the live environments continue from where the rollout stopped
    o2, l2 = env.engine.observe(pmax)
    assert int((l2 > 0).sum()) == N
    # determinism: same seeds, same counters -> identical trajectories
    env2 = LeadMonomialsEnv("3-20-10-weighted", k=2, num_envs=N, pmax=pmax)
    env2.seed(np.arange(10, 10 + N))
    env2.reset()
    tb2 = collect(env2, net, T, counter0=1000, store_obs=False)
    assert torch.equal(tb2.actions, tb.actions) and torch.equal(tb2.reward, tb.reward)
