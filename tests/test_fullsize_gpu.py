"""BASELINE.json's FULL sizes through size-independent properties (the oracle needs hours there): the replay minibatch of
4096 graphs x 200 vertices (configs[4]) and the 10 000 x 50-vertex rollout sweep (configs[2]).

fwd+bwd: the minibatch mean is a sum over independent graphs, so loss and gradient must not depend on HOW the minibatch is
cut into micro-batches (512- vs 384-graph cuts: different tile alignments, a ragged last micro-batch) nor on the ORDER of
the graphs (a permuted minibatch gives the same loss and gradient), and a doubled minibatch (every transition twice) has the
same mean gradient -- each to fp32 summation-order tolerance; the small-size tests pin the same code against the oracle.
Rollouts: every colouring is proper and complete, uses max colour + 1 colours, obeys the first-fit bound (at most
max degree + 1 colours: first fit never needs more), and the CUDA-graph replay gives the eager result bit for bit.
GCRL_FULLSIZE_GRAPHS / GCRL_FULLSIZE_ROLLOUT_GRAPHS shrink the sizes for a quick run."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
G = int(os.environ.get('GCRL_FULLSIZE_GRAPHS', '4096'))
GR = int(os.environ.get('GCRL_FULLSIZE_ROLLOUT_GRAPHS', '10000'))


def er_batch(n_graphs, n, p, seed):
    rng = np.random.default_rng(seed)
    out = np.zeros((n_graphs, n, n), dtype=np.uint8)
    iu = np.triu_indices(n, 1)
    for g in range(n_graphs):
        out[g][iu] = rng.random(iu[0].size) < p
        out[g] |= out[g].T
    return out


def test_full_size_minibatch_is_invariant_to_micro_batching_order_and_duplication():
    if torch.cuda.get_device_properties(0).total_memory < 120 * 2 ** 30:
        pytest.skip('needs a 180 GB B200')
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    from graph_colouring_with_rl_b200.trainer import FIELDS, QLearner, TransitionBatch, shard_host, synthetic_transitions
    n = 200
    parts = []
    for c in range((G + 511) // 512):
        k = min(512, G - 512 * c)
        parts.append(synthetic_transitions(list(er_batch(k, n, 0.5, 1000 + c)), 0.5, seed=1 + c, device=DEV))
    ns = np.concatenate([p[0] for p in parts])
    host = {k: torch.cat([p[1][k] for p in parts]) for k in FIELDS}
    assert int(host['action'].min()) >= 0
    torch.manual_seed(0)
    agent = DQNAgentGN(2, 1, 0, math_mode='bf16x3')

    def run(ns_, host_, micro_graphs):
        learner = QLearner(agent, max_edges_per_micro_batch=micro_graphs * n * (n - 1), local_only=True)
        batch = TransitionBatch.from_host(ns_, host_, DEV)
        q_next = learner.target_q(batch)
        loss = float(learner.fwd_bwd(batch, q_next).item())
        g = agent.gn_behaviour._flat_grad.double().clone()
        del batch, learner
        torch.cuda.empty_cache()
        return loss, g

    loss_a, g_a = run(ns, host, 512)
    assert np.isfinite(loss_a) and bool(torch.isfinite(g_a).all()) and float(g_a.norm()) > 0
    scale = float(g_a.abs().max())
    # (1) another micro-batch cut (ragged last one, other tile alignments)
    loss_b, g_b = run(ns, host, 384)
    assert abs(loss_b - loss_a) <= 1e-5 * abs(loss_a), (loss_a, loss_b)
    assert float((g_b - g_a).abs().max()) <= 2e-4 * scale
    # (2) the same transitions in another order
    perm = np.random.default_rng(5).permutation(G)
    pieces = [shard_host(ns, host, int(i), int(i) + 1) for i in perm]
    ns_p = np.concatenate([p[0] for p in pieces])
    host_p = {k: torch.cat([p[1][k] for p in pieces]) for k in FIELDS}
    loss_c, g_c = run(ns_p, host_p, 512)
    assert abs(loss_c - loss_a) <= 1e-5 * abs(loss_a), (loss_a, loss_c)
    assert float((g_c - g_a).abs().max()) <= 2e-4 * scale
    # (3) every transition twice: the MEAN loss and gradient do not move (first half of the minibatch to keep it in memory)
    h = G // 2
    ns_h, host_h = shard_host(ns, host, 0, h)
    loss_h, g_h = run(ns_h, host_h, 512)
    ns_d = np.concatenate([ns_h, ns_h])
    host_d = {k: torch.cat([host_h[k], host_h[k]]) for k in FIELDS}
    loss_d, g_d = run(ns_d, host_d, 512)
    assert abs(loss_d - loss_h) <= 1e-5 * abs(loss_h), (loss_h, loss_d)
    assert float((g_d - g_h).abs().max()) <= 2e-4 * float(g_h.abs().max())


def test_full_size_rollout_sweep_properties():
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
    from graph_colouring_with_rl_b200.rollouts import greedy_rollouts
    n = 50
    adjs = er_batch(GR, n, 0.5, 0)
    torch.manual_seed(0)
    agent = DQNAgentGN(2, 1, 0, test_mode=True, math_mode='bf16x3')
    env = BatchedColouringEnv(list(adjs), device=DEV)
    res = greedy_rollouts(None, agent.gn_behaviour, env=env, graph=True)
    col = env.colour.cpu().numpy().reshape(GR, n)
    used = res['colours_used'].cpu().numpy()
    assert col.min() >= 0, 'an episode did not finish'
    A = adjs.astype(bool)
    same = col[:, :, None] == col[:, None, :]
    assert not np.any(A & same), 'two adjacent vertices share a colour'
    assert np.array_equal(used, col.max(1) + 1)
    assert np.all(used <= adjs.sum(2).max(1) + 1), 'first fit used more than max degree + 1 colours'
    assert bool(env.done.all().item())
    # every colour class below the maximum is inhabited (first fit never skips a colour)
    for g in np.random.default_rng(1).choice(GR, size=min(GR, 200), replace=False):
        assert len(set(col[g].tolist())) == used[g]
    eager = greedy_rollouts(None, agent.gn_behaviour, env=env, graph=False)
    assert torch.equal(eager['colours_used'], res['colours_used']) and eager['steps'] == res['steps']
    assert np.array_equal(env.colour.cpu().numpy().reshape(GR, n), col)
