"""GPU parity tests on the graph shapes of BASELINE.json configs A and D: the reconstructible part of
the Lemos benchmark (queen / Mycielski graphs) and single large graphs of 100-1000
vertices, against the CPU oracle (oracle/gn_oracle.py + oracle/env_oracle.py) on the parity weights."""
import os

import numpy as np
import pytest
import torch

import gcrl_testutil as util
from oracle import env_oracle, gn_oracle

pytestmark = pytest.mark.gpu
WEIGHTS_ROOT = os.path.join(util.GOLD, 'parity_weights')


def make_agent(math_mode='fp32', root=WEIGHTS_ROOT):
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    torch.manual_seed(0)
    agent = DQNAgentGN(2, 1, 0, test_mode=True, math_mode=math_mode)
    agent.load_models(root)
    return agent


def load_weights(which):
    """'parity': the reference agent's own state_dict after 160 CPU episodes; 'trained25k': the behaviour network after the
    reference's whole 25 000-episode schedule run by the device training loop (scripts/train_full_schedule.py)."""
    if which == 'parity':
        return WEIGHTS_ROOT, util.load_parity_weights()
    from collections import OrderedDict
    root = os.path.join(util.GOLD, 'trained25k')
    sd = torch.load(root + '_GN', map_location='cpu')
    return root, OrderedDict((k, v.float()) for k, v in sd.items())


def oracle_q(params, adj, colour):
    x, ei, ea = gn_oracle.observation(adj, colour)
    with torch.no_grad():
        _, q = gn_oracle.gn_forward(params, x, ei, ea)
    return q.view(-1).numpy()


@pytest.mark.parametrize('weights', ['parity', 'trained25k'])
@pytest.mark.parametrize('math_mode,tol', [('fp32', 5e-5), ('bf16x3', 1e-4)])
def test_lemos_reconstructible_rollouts_vs_oracle(math_mode, tol, weights):
    """Greedy rollouts (first-vertex rule, argmax Q, first fit, leaf pass) on queen / Mycielski graphs:
    the device sequence equals the oracle's up to the first decision whose top-2 Q gap is below 4*tol;
    every colouring is proper and complete."""
    from graph_colouring_with_rl_b200.dataset_functions import (MYCIEL_ORDER, QUEEN_BOARDS, lemos_reconstructible,
                                                                run_learned_policy_on_dataset)
    from graph_colouring_with_rl_b200.rollouts import greedy_rollouts, sequences_from_actions
    names = ['queen5_5', 'queen6_6', 'queen7_7', 'queen8_8', 'queen8_12', 'queen10_10', 'myciel5', 'myciel6']
    graphs = lemos_reconstructible(names)
    root, params = load_weights(weights)
    agent = make_agent(math_mode, root)
    if weights == 'trained25k' and math_mode == 'bf16x3':
        tol = 3e-4                          # |dQ| of bf16x3 on fully trained weights (|Q| up to ~10 on the 100-vertex boards)
    res = greedy_rollouts([a for _, a in graphs], agent.gn_behaviour, record=True)
    seqs = sequences_from_actions(res['actions'])
    used = res['colours_used'].cpu().numpy()
    col = res['env'].colour.cpu().numpy()
    off = res['env'].node_off.cpu().numpy()
    full = checked = 0
    for i, (nm, adj) in enumerate(graphs):
        cu, seq, _, gaps = env_oracle.greedy_rollout(adj, lambda c, a=adj: oracle_q(params, a, c))
        gaps = np.asarray(gaps)
        if gaps.size == 0 or gaps.min() > 4 * tol:
            assert seqs[i] == seq and used[i] == cu, nm
            full += 1
            checked += len(seq)
        else:
            k = int(np.argmax(gaps <= 4 * tol)) + 1
            assert seqs[i][:k] == seq[:k], nm
            checked += k
        c = col[off[i]:off[i + 1]]
        assert c.min() >= 0 and not np.any(adj.astype(bool) & (c[:, None] == c[None, :])), nm
        assert used[i] == c.max() + 1
        # lower bounds: a queen graph on an r x c board holds a max(r, c)-clique; M_k has chromatic number k
        lower = max(QUEEN_BOARDS[nm]) if nm in QUEEN_BOARDS else MYCIEL_ORDER[nm]
        assert used[i] >= lower, nm
    # the boards are highly symmetric, so near-ties are common: count the decisions that could be compared
    print('%s / %s weights: %d graphs compared to the end, %d decisions compared in total; colours used %s'
          % (math_mode, weights, full, checked, dict(zip(names, used.tolist()))))
    assert full >= 1 and checked >= 40
    targets = lemos_reconstructible(names[:3], as_targets=True)
    stats, by = run_learned_policy_on_dataset(targets, agent)
    assert [b['name'] for b in by] == names[:3] and by[0]['min_colours_used'] == used[0]


@pytest.mark.parametrize('n,p', [(100, 0.5), (250, 0.1), (500, 0.9), (1000, 0.5)])
def test_large_single_graph_forward_vs_oracle(n, p):
    """Config D: one graph of 100-1000 vertices (up to 999 000 complete-graph edges) through the
    reference-facing GNNetwork.forward, mid-episode state; Q against the CPU oracle."""
    from graph_colouring_with_rl_b200.architecture import GNNetwork
    from graph_colouring_with_rl_b200.data import Data
    rng = np.random.default_rng(n)
    adj = env_oracle.er_graph(n, p, rng)
    col = util.random_colours(n, adj, 0.4, rng)
    params = util.load_parity_weights()
    x, ei, ea = gn_oracle.observation(adj, col)
    with torch.no_grad():
        h_ref, q_ref = gn_oracle.gn_forward(params, x, ei, ea)
    q_ref = q_ref.view(-1).numpy()
    for mode, tol in (('fp32', 1e-4), ('bf16x3', 1e-3), ('bf16', 0.15)):
        net = GNNetwork(2, 1, 0, name='GN', alpha=1e-3, math_mode=mode)
        net.load_checkpoint(WEIGHTS_ROOT)
        data = Data(x=x.clone(), edge_index=ei.clone(), edge_attr=ea.clone()).to(net.device)
        with torch.no_grad():
            h, q = net(data, torch.zeros(1, 0, device=net.device), [0] * n, None)
        q = q.view(-1).cpu().numpy()
        err = np.abs(q - q_ref).max()
        print('n=%d p=%.1f %s: max |dQ| = %.2e (|Q| max %.2f)' % (n, p, mode, err, np.abs(q_ref).max()))
        assert err < tol * max(1.0, np.abs(q_ref).max()), (mode, err)
        if mode == 'fp32':
            np.testing.assert_allclose(h.cpu().numpy(), h_ref.numpy(), atol=2e-4, rtol=1e-4)
    # the argmax the agent would take agrees with the oracle's when the gap allows
    un = np.nonzero(col < 0)[0]
    srt = np.sort(q_ref[un])
    if srt[-1] - srt[-2] > 2e-3:
        assert un[np.argmax(q[un])] == un[np.argmax(q_ref[un])]
