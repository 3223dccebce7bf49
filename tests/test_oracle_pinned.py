"""CPU tests that pin oracle/ (the CPU restatement) against the golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py -> tests/golden/).  `not gpu`."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import env_oracle, gn_oracle
import gcrl_testutil as util

GRAPHS = util.load_validation_graphs()


# ---- environment (integer, exact) ----------------------------------------------------------
def test_dsatur_known_answers_all_100():
    """Stored DSATUR vertex_order -> colours_used (validation_dataset.pickle)."""
    for g in GRAPHS:
        used, done, col = env_oracle.replay_order(g['adj'], g['dsatur_order'])
        assert done and used == g['dsatur_colours'], g['name']
        a = g['adj'].astype(bool)
        assert not np.any(a & (col[:, None] == col[None, :]))       # proper colouring


def test_env_traces_step_by_step():
    d = np.load(util.GOLD + '/env_traces.npz')
    acts = util.split_by(d['actions'], d['steps'])
    rews = util.split_by(d['rewards'], d['steps'])
    dones = util.split_by(d['dones'], d['steps'])
    co = 0
    for gi, a_seq, r_seq, d_seq in zip(d['graph_ids'], acts, rews, dones):
        g = GRAPHS[int(gi)]
        env = env_oracle.ColouringEnvOracle(g['adj'])
        for a, r, dn in zip(a_seq, r_seq, d_seq):
            col, rew, done = env.step(int(a))
            ref_col = d['colours'][co:co + g['n']].astype(np.int64)
            co += g['n']
            assert np.array_equal(col, ref_col)
            assert rew == int(r) and int(done) == int(dn)
    assert co == d['colours'].shape[0]


def test_neighbour_dist_counts_and_first_vertex():
    for g in GRAPHS:
        counts = env_oracle.neighbour_dist_counts(g['adj'])
        assert np.array_equal(counts.astype(np.int64), g['dist_counts']), g['name']
    roll = np.load(util.GOLD + '/rollouts.npz')
    seqs = util.split_by(roll['seq'], roll['seq_len'])
    for g, s in zip(GRAPHS, seqs):
        assert env_oracle.first_vertex(g['adj']) == int(s[0])


def test_edge_order_fixture_is_complete():
    for g in GRAPHS:
        n, el = g['n'], g['edge_list']
        assert el.shape[0] == n * (n - 1)
        assert len(set(map(tuple, el.tolist()))) == n * (n - 1)


# ---- network (fp32 differential vs the reference's own modules) -------------------------------
def _states(colours, gids):
    out, o = [], 0
    for gi in gids:
        g = GRAPHS[int(gi)]
        col = colours[o:o + g['n']].astype(np.int64)
        o += g['n']
        out.append((g, col))
    return out


def test_default_init_matches_reference_constructor():
    d = np.load(util.GOLD + '/gn_forward.npz')
    p = gn_oracle.init_params(seed=0)
    assert list(p.keys()) == [str(k) for k in d['keys']]
    sums = np.asarray([float(v.double().sum()) for v in p.values()])
    abss = np.asarray([float(v.double().abs().sum()) for v in p.values()])
    np.testing.assert_allclose(sums, d['init_param_sum'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(abss, d['init_param_abs'], rtol=0, atol=1e-9)
    assert sum(v.numel() for v in p.values()) == 198529


@pytest.mark.parametrize('tag', ['init', 'trained'])
def test_forward_matches_reference(tag):
    d = np.load(util.GOLD + '/gn_forward.npz')
    p = gn_oracle.init_params(seed=0) if tag == 'init' else util.load_parity_weights()
    qs, hs, acts = [], [], []
    for g, col in _states(d['colours'], d['graph_ids']):
        x, ei, ea = gn_oracle.observation(g['adj'], col, g['edge_list'])
        with torch.no_grad():
            h, q = gn_oracle.gn_forward(p, x, ei, ea)
        qs.append(q.view(-1).numpy())
        hs.append(h.double().sum(1).numpy())
        acts.append(gn_oracle.greedy_action(q.numpy(), [int(v) for v in np.nonzero(col < 0)[0]]))
    q = np.concatenate(qs)
    tol = 1e-6 if tag == 'trained' else 2e-8
    np.testing.assert_allclose(q, d[tag + '_q'], rtol=0, atol=tol)
    np.testing.assert_allclose(np.concatenate(hs), d[tag + '_h5_rowsum'], rtol=1e-5, atol=1e-5)
    if tag == 'trained':
        assert np.array_equal(np.asarray(acts), d[tag + '_actions'])


def test_learn_step_matches_reference():
    d = np.load(util.GOLD + '/learn_step.npz')
    pb = util.load_parity_weights()
    pt = util.unflatten(torch.from_numpy(d['target_before']), pb)
    cur = [gn_oracle.observation(g['adj'], c, g['edge_list']) for g, c in _states(d['cur_colours'], d['graph_ids'])]
    nxt_states = _states(d['next_colours'], d['graph_ids'])
    nxt = [gn_oracle.observation(g['adj'], c, g['edge_list']) for g, c in nxt_states]
    counts = [g['n'] for g, _ in nxt_states]
    assigned = np.concatenate([(c >= 0).astype(np.int64) for _, c in nxt_states])
    res = gn_oracle.learn_step(pb, pt, {}, gn_oracle.collate(cur), gn_oracle.collate(nxt), counts, d['actions'],
                               d['rewards'], d['dones'], assigned)
    assert abs(res['loss'] - float(d['loss'])) < 1e-6
    np.testing.assert_allclose(res['q_expected'].view(-1).numpy(), d['q_expected'], atol=1e-6)
    np.testing.assert_allclose(res['q_targets'].view(-1).numpy(), d['q_targets'], atol=1e-6)
    g = util.flatten(res['grads']).numpy()
    np.testing.assert_allclose(g, d['grads'], rtol=1e-4, atol=1e-6)
    # Adam's first step moves every touched weight by ~lr; sign flips can only occur where |g| ~ 0
    after = util.flatten(res['params_b']).numpy()
    diff = np.abs(after - d['behaviour_after'])
    assert np.mean(diff > 1e-6) < 1e-3
    np.testing.assert_allclose(util.flatten(res['params_t']).numpy(), d['target_after'], atol=3e-6)


def test_rollouts_match_reference():
    d = np.load(util.GOLD + '/rollouts.npz')
    p = util.load_parity_weights()
    seqs = util.split_by(d['seq'], d['seq_len'])
    gaps = util.split_by(d['gaps'], d['gaps_len'])
    for gi in range(0, 100, 9):
        g = GRAPHS[gi]

        def q_fn(col, g=g):
            x, ei, ea = gn_oracle.observation(g['adj'], col, g['edge_list'])
            with torch.no_grad():
                return gn_oracle.gn_forward(p, x, ei, ea)[1].view(-1).numpy()
        used, seq, col, og = env_oracle.greedy_rollout(g['adj'], q_fn, g['edge_list'])
        if gaps[gi].size == 0 or gaps[gi].min() > 1e-5:
            assert seq == [int(v) for v in seqs[gi]], g['name']
            assert used == int(d['colours_used'][gi])


def test_rollouts_with_fully_trained_weights_match_reference():
    """The same pin on the FULLY trained policy (tests/golden/trained25k_GN: the reference's 25 000-episode schedule run by the
    device training loop; rollouts_trained25k.npz: what the unmodified reference does with it)."""
    import os
    from collections import OrderedDict
    d = np.load(util.GOLD + '/rollouts_trained25k.npz')
    sd = torch.load(os.path.join(util.GOLD, 'trained25k_GN'), map_location='cpu')
    p = OrderedDict((k, v.float()) for k, v in sd.items())
    seqs = util.split_by(d['seq'], d['seq_len'])
    gaps = util.split_by(d['gaps'], d['gaps_len'])
    checked = 0
    for gi in range(2, 100, 8):
        g = GRAPHS[gi]

        def q_fn(col, g=g):
            x, ei, ea = gn_oracle.observation(g['adj'], col, g['edge_list'])
            with torch.no_grad():
                return gn_oracle.gn_forward(p, x, ei, ea)[1].view(-1).numpy()
        used, seq, col, og = env_oracle.greedy_rollout(g['adj'], q_fn, g['edge_list'])
        if gaps[gi].size == 0 or gaps[gi].min() > 1e-5:
            assert seq == [int(v) for v in seqs[gi]], g['name']
            assert used == int(d['colours_used'][gi])
            checked += 1
    assert checked >= 8
    assert abs(float(d['colours_used'].mean()) - 7.45) < 1e-9      # (DSATUR on the same graphs: 7.42, random order: 8.58)


def test_replay_oracle_matches_reference_replay_buffer():
    """oracle/replay_oracle.py against the reference ReplayBuffer (tests/golden/replay_buffer.npz): ring index, wrap-around
    and the minibatches three seeded sample_buffer calls returned."""
    from oracle.replay_oracle import ReplayOracle
    d = np.load(util.GOLD + '/replay_buffer.npz')
    graphs = util.load_validation_graphs()
    cap = int(d['capacity'])
    orc = ReplayOracle(cap, [g['adj'] for g in graphs])
    ns = [graphs[int(g)]['n'] for g in d['graph_ids']]
    cur, nxt = util.split_by(d['cur'], ns), util.split_by(d['nxt'], ns)
    for i, g in enumerate(d['graph_ids']):
        slot = orc.store(int(g), cur[i], nxt[i], d['actions'][i], d['rewards'][i], d['dones'][i])
        assert slot == i % cap                                              # dqn_agent.py:33
    assert orc.count == int(d['count'])
    np.random.seed(23)
    for b in range(3):
        idx = np.random.choice(min(orc.count, cap), 6)                      # dqn_agent.py:47-48
        out = orc.gather(idx)
        assert np.array_equal(out['ns'], d['s%d_nodes' % b])
        assert np.array_equal(out['cur_colour'], d['s%d_x' % b]) and np.array_equal(out['next_colour'], d['s%d_nx' % b])
        assert np.array_equal(out['action'], d['s%d_actions' % b]) and np.array_equal(out['done'], d['s%d_dones' % b])
        assert np.array_equal(out['reward'], d['s%d_rewards' % b])
        assert np.array_equal((out['next_colour'] >= 0).astype(np.int8), d['s%d_assigned' % b])
        assert int(out['adj'].sum()) == int(d['s%d_present_edges' % b])     # both directions of every edge, like edge_attr == -1
