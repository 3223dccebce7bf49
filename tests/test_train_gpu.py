"""GPU tests of the device replay ring and the device training loop:
gcrl_replay_store / gcrl_replay_gather bit-exact against oracle/replay_oracle.py, and
dqn_main.DeviceTrainer against a golden run of the reference's own training loop
(tests/golden/train_loop.npz, oracle/make_golden.py stage_trainloop)."""
import os
import random

import numpy as np
import pytest
import torch

import gcrl_testutil as util
from oracle import env_oracle
from oracle.replay_oracle import ReplayOracle

pytestmark = pytest.mark.gpu
GRAPHS = util.load_validation_graphs()
WEIGHTS_ROOT = os.path.join(util.GOLD, 'parity_weights')


def make_agent(trained=True):
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    torch.manual_seed(0)
    agent = DQNAgentGN(2, 1, 0)
    if trained:
        agent.load_models(WEIGHTS_ROOT)
        agent.soft_update(agent.gn_behaviour, agent.gn_target, 1)
    return agent


def _play_and_store(adjs, capacity, n_steps, seed, stride=None):
    """Random episodes on a batched env; every step's transitions go to the device ring and to the oracle."""
    from graph_colouring_with_rl_b200 import ops
    from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
    from graph_colouring_with_rl_b200.replay import DeviceReplayBuffer, GraphPool
    pool = GraphPool(adjs)
    buf = DeviceReplayBuffer(capacity, pool, stride=stride)
    orc = ReplayOracle(capacity, adjs)
    rng = np.random.default_rng(seed)
    ids = list(rng.permutation(len(adjs)))
    env = BatchedColouringEnv.from_pool(pool, ids)
    off = env.node_off.cpu().numpy()
    gen = torch.Generator(device=env.device)
    gen.manual_seed(seed)
    for step in range(n_steps):
        r = torch.rand(env.n_vertices, generator=gen, device=env.device)
        a = ops.score_argmax(r, env.colour, env.node_off)
        mask = rng.random(len(ids)) < 0.7
        done_before = env.done.cpu().numpy().astype(bool)
        mask &= ~done_before
        cur = env.colour.clone()
        env.step(a)
        slots = buf.store_from_env(env, ids, cur, a, mask=mask)
        cur_h, nxt_h = cur.cpu().numpy(), env.colour.cpu().numpy()
        a_h, r_h, d_h = a.cpu().numpy(), env.reward.cpu().numpy(), env.done.cpu().numpy()
        k = 0
        for e in np.nonzero(mask)[0]:
            s = orc.store(ids[e], cur_h[off[e]:off[e + 1]], nxt_h[off[e]:off[e + 1]], a_h[e], r_h[e], d_h[e])
            assert s == slots[k]
            k += 1
    return pool, buf, orc


def _check_gather(buf, orc, batch):
    tb = buf.gather(batch)
    buf.check()
    ref = orc.gather(batch)
    assert np.array_equal(tb.ns, ref['ns'])
    assert np.array_equal(tb.adj.cpu().numpy(), ref['adj'])
    for k in ('cur_colour', 'next_colour', 'action', 'done'):
        assert np.array_equal(getattr(tb, k).cpu().numpy(), ref[k]), k
    assert np.array_equal(tb.reward.cpu().numpy(), ref['reward'])
    return tb


def test_replay_ring_store_gather_bit_exact_with_wraparound():
    adjs = [g['adj'] for g in GRAPHS[:24]]
    pool, buf, orc = _play_and_store(adjs, capacity=97, n_steps=12, seed=5)
    assert buf.current_mem_size == orc.count > 97                       # the ring wrapped
    rng = np.random.default_rng(1)
    _check_gather(buf, orc, rng.integers(0, 97, size=64))              # repeats allowed (sampling with replacement)
    _check_gather(buf, orc, np.arange(97))
    _check_gather(buf, orc, np.asarray([3]))
    np.random.seed(11)
    idx = buf.sample_indices(16)
    np.random.seed(11)
    assert np.array_equal(idx, np.random.choice(97, 16))               # dqn_agent.py:47-48


def test_device_ring_reproduces_reference_replay_buffer_golden():
    """The reference ReplayBuffer's own run (tests/golden/replay_buffer.npz: 29 transitions into 11 slots, three seeded
    sample_buffer calls) through gcrl_replay_store / gcrl_replay_gather."""
    from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
    from graph_colouring_with_rl_b200.replay import DeviceReplayBuffer, GraphPool
    d = np.load(util.GOLD + '/replay_buffer.npz')
    pool = GraphPool([g['adj'] for g in GRAPHS])
    buf = DeviceReplayBuffer(int(d['capacity']), pool)
    ns = [GRAPHS[int(g)]['n'] for g in d['graph_ids']]
    cur, nxt = util.split_by(d['cur'], ns), util.split_by(d['nxt'], ns)
    dev = pool.device
    for i, g in enumerate(d['graph_ids']):
        env = BatchedColouringEnv.from_pool(pool, [int(g)])
        env.colour.copy_(torch.from_numpy(nxt[i].astype(np.int32)).to(dev))
        env.reward.fill_(float(d['rewards'][i]))
        env.done.fill_(int(d['dones'][i]))
        buf.store_from_env(env, [int(g)], torch.from_numpy(cur[i].astype(np.int32)).to(dev),
                           torch.tensor([int(d['actions'][i])], dtype=torch.int32, device=dev))
    assert buf.current_mem_size == int(d['count'])
    np.random.seed(23)
    for b in range(3):
        tb = buf.sample_buffer(6)                                            # np.random.choice like dqn_agent.py:47-48
        buf.check()
        assert np.array_equal(tb.ns, d['s%d_nodes' % b])
        assert np.array_equal(tb.cur_colour.cpu().numpy(), d['s%d_x' % b])
        assert np.array_equal(tb.next_colour.cpu().numpy(), d['s%d_nx' % b])
        assert np.array_equal(tb.action.cpu().numpy(), d['s%d_actions' % b])
        assert np.array_equal(tb.reward.cpu().numpy(), d['s%d_rewards' % b])
        assert np.array_equal(tb.done.cpu().numpy(), d['s%d_dones' % b])
        assert int(tb.adj.sum().item()) == int(d['s%d_present_edges' % b])


def test_replay_gather_vector_path_and_padding():
    # vertex counts that are multiples of 4 keep every adjacency block 16-byte aligned (uint4 copies);
    # a stride larger than the largest graph exercises the -1 padding
    rng = np.random.default_rng(3)
    adjs = [env_oracle.er_graph(n, 0.5, rng) for n in (64, 200, 132, 16, 8, 148)]
    pool, buf, orc = _play_and_store(adjs, capacity=40, n_steps=6, seed=9, stride=256)
    n_written = min(buf.current_mem_size, 40)
    tb = _check_gather(buf, orc, np.random.default_rng(0).integers(0, n_written, size=33))
    r_cur = buf.ring[1].cpu().numpy()
    s = 0
    n0 = pool.ns[buf.slot_graph_h[s]]
    assert (r_cur[s, n0:] == -1).all()
    assert tb.n_edges == int((tb.ns * (tb.ns - 1)).sum())


def test_replay_gather_reports_stale_host_mirror():
    adjs = [g['adj'] for g in GRAPHS[:6]]
    pool, buf, orc = _play_and_store(adjs, capacity=32, n_steps=3, seed=2)
    buf.gather(np.asarray([0, 1]))
    buf.check()
    s = 0
    other = next(i for i in range(6) if pool.ns[i] != pool.ns[buf.slot_graph_h[s]])
    buf.slot_graph_h[s] = other                                          # host mirror no longer matches the ring
    buf.gather(np.asarray([1, 0]))
    with pytest.raises(RuntimeError):
        buf.check()
    with pytest.raises(IndexError):
        buf.gather(np.asarray([31]))                                     # never written


def test_learn_from_device_ring_matches_reference_golden():
    """The reference's learn() golden (8 transitions) replayed through ring -> gather -> QLearner."""
    from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
    from graph_colouring_with_rl_b200.replay import DeviceReplayBuffer, GraphPool
    from graph_colouring_with_rl_b200.trainer import QLearner
    d = np.load(util.GOLD + '/learn_step.npz')
    agent = make_agent()
    agent.gn_target.flat_params().copy_(torch.from_numpy(d['target_before']).to(agent.gn_target.device))
    gids = [int(g) for g in d['graph_ids']]
    pool = GraphPool([GRAPHS[g]['adj'] for g in gids])
    buf = DeviceReplayBuffer(16, pool)
    env = BatchedColouringEnv.from_pool(pool, list(range(8)))
    ns = [GRAPHS[g]['n'] for g in gids]
    cur = torch.from_numpy(d['cur_colours'].astype(np.int32)).to(env.device)
    env.colour.copy_(torch.from_numpy(d['next_colours'].astype(np.int32)).to(env.device))
    env.reward.copy_(torch.from_numpy(d['rewards']).to(env.device))
    env.done.copy_(torch.from_numpy(d['dones']).to(env.device))
    act = torch.from_numpy(d['actions']).to(env.device)
    buf.store_from_env(env, list(range(8)), cur, act)
    batch = buf.gather(np.arange(8))
    assert sum(ns) == batch.n_vertices
    loss = QLearner(agent).learn_step(batch)
    assert abs(float(loss.item()) - float(d['loss'])) < 1e-5
    diff = np.abs(agent.gn_behaviour.flat_params().cpu().numpy() - d['behaviour_after'])
    assert np.mean(diff > 1e-6) < 1e-3
    np.testing.assert_allclose(agent.gn_target.flat_params().cpu().numpy(), d['target_after'], atol=3e-6)


def test_training_loop_follows_the_reference_run():
    """n_envs=1 + host RNG consumes random / np.random in the reference's order: with the same seeds,
    weights and graphs the run must take the reference's episodes step for step.  fp32 rounding can only
    change an action where the reference's top-2 Q gap is tiny; everything is compared up to the first
    such step (if any), and that step must be NN-chosen with a gap below GAP_TOL."""
    from graph_colouring_with_rl_b200.dqn_main import DeviceTrainer
    GAP_TOL = 2e-3
    d = np.load(util.GOLD + '/train_loop.npz')
    agent = make_agent()
    agent.batch_size = 8
    agent.no_episodes = 60
    random.seed(3)
    np.random.seed(3)
    trace = {}
    tr = DeviceTrainer(agent, [g['adj'] for g in GRAPHS], None, n_envs=1, host_rng=True, trace=trace)
    tr.run(len(d['graph_ids']), verbose=False)
    ref_acts = util.split_by(d['actions'], d['ep_len'])
    ref_nn = util.split_by(d['nn_chosen'], d['ep_len'])
    ref_gap = util.split_by(d['gaps'], d['ep_len'])
    steps_ok, diverged = 0, None
    for ep, gi in enumerate(d['graph_ids']):
        if trace['graph_ids'][ep] != int(gi):
            diverged = diverged or (ep, -1)
            break
        mine = trace['actions'][ep]
        for t, a in enumerate(ref_acts[ep]):
            if t >= len(mine) or mine[t] != int(a):
                diverged = (ep, t)
                break
            steps_ok += 1
        if diverged:
            break
        assert len(mine) == len(ref_acts[ep])
        assert tr.training_stats.colours_used[ep] == int(d['colours_used'][ep])
        assert tr.training_stats.episode_rewards[ep] == int(d['scores'][ep])
    losses = np.asarray([float(l.item()) for l in tr.losses])
    if diverged is None:
        assert steps_ok == int(d['ep_len'].sum())
        assert tr.memory.current_mem_size == int(d['buffer_size'])
        assert len(losses) == len(d['losses'])
        n_cmp = len(losses)
    else:
        ep, t = diverged
        assert t >= 0 and ref_nn[ep][t] == 1 and ref_gap[ep][t] < GAP_TOL, (diverged, float(ref_gap[ep][t]))
        assert steps_ok >= 150                                        # well past the point where learning starts
        first_step = int(d['ep_len'][:ep].sum()) + t                  # global step index of the divergence
        n_cmp = max(0, min(len(losses), (first_step - 1) // agent.update_every - 20))
    print('train loop parity: %d/%d steps identical, %d losses compared, divergence at %s'
          % (steps_ok, int(d['ep_len'].sum()), n_cmp, diverged))
    assert n_cmp >= 10
    # losses: same minibatches (np.random stream), weights drift apart only by fp32 rounding through Adam
    np.testing.assert_allclose(losses[:n_cmp], d['losses'][:n_cmp], rtol=5e-2, atol=2e-3)
    np.testing.assert_allclose(losses[:3], d['losses'][:3], rtol=1e-3, atol=1e-5)


def test_batched_training_waves_cadence_validation_and_files(tmp_path):
    from graph_colouring_with_rl_b200.dqn_main import DeviceTrainer
    from graph_colouring_with_rl_b200 import utils as U
    agent = make_agent()
    agent.batch_size = 16
    agent.no_episodes = 40
    w0 = agent.gn_behaviour.flat_params().clone()
    train_graphs = [g['adj'] for g in GRAPHS[:40]]
    val_graphs = [dict(name=g['name'], adj=g['adj'], no_vertices=g['n']) for g in GRAPHS[40:60]]
    trace = {}
    tr = DeviceTrainer(agent, train_graphs, val_graphs, n_envs=12, validate_every=24, out_dir=str(tmp_path),
                       device_seed=1, trace=trace)
    tr.run(40, verbose=False)
    assert tr.episodes_done == 40 and agent.episode_no == 40
    n_steps = sum(len(s) for s in trace['actions'])
    assert tr.t_step == n_steps
    assert tr.memory.current_mem_size == n_steps - 40                   # first steps are not stored (dqn_main.py:182)
    # every episode is a valid complete colouring of its graph, replayed through the oracle env
    for gi, seq, used, score in zip(trace['graph_ids'], trace['actions'], tr.training_stats.colours_used,
                                    tr.training_stats.episode_rewards):
        cu, done, _ = env_oracle.replay_order(train_graphs[gi], seq)
        assert done and cu == used and score == -used
    # learn cadence: one learn per update_every env steps once 10 x batch transitions are stored
    assert 0 < len(tr.losses) <= n_steps // agent.update_every + 1
    assert not torch.equal(w0, agent.gn_behaviour.flat_params())
    assert all(np.isfinite(float(l.item())) for l in tr.losses)
    # stored transitions are consistent: next colours = oracle env step of cur colours with the stored action
    tb = tr.memory.gather(np.arange(min(tr.memory.current_mem_size, 64)))
    tr.memory.check()
    off = tb.node_off_h
    cur, nxt, act = tb.cur_colour.cpu().numpy(), tb.next_colour.cpu().numpy(), tb.action.cpu().numpy()
    rew, done = tb.reward.cpu().numpy(), tb.done.cpu().numpy()
    for b in range(tb.n_graphs):
        gi = int(tr.memory.slot_graph_h[b])
        e = env_oracle.ColouringEnvOracle(train_graphs[gi])
        e.colour = cur[off[b]:off[b + 1]].astype(np.int64).copy()
        e.max_colour = int(e.colour.max())
        col, r, dn = e.step(int(act[b]))
        assert np.array_equal(col, nxt[off[b]:off[b + 1]]) and r == rew[b] and int(dn) == done[b]
    # validation sweeps at 0, 24, 40(no: 40 % 24 != 0) -> keys 0 and 24; best checkpoint kept, older one unlinked
    assert sorted(tr.all_validation_stats) == [0, 24]
    assert len(tr.all_validation_stats[24].colours_used) == 20
    params = sorted(os.listdir(tr.params_dir))
    assert 'final_policy_episode_40_GN' in params
    best = [p for p in params if p.startswith('best_policy_episode_')]
    assert len(best) == 1 and os.path.basename(tr.best_validation_filepath) == best[0]
    files = sorted(os.listdir(tr.stats_dir))
    assert files == ['training_stats_episode_40.txt', 'validation_stats_episode_40.txt']
    st = U.load_stats_from_file(os.path.join(tr.stats_dir, files[0]))
    assert st.saved_episodes == [float(i) for i in range(1, 41)]
    assert st.colours_used == [float(c) for c in tr.training_stats.colours_used]
    vs = U.load_validation_stats_from_file(os.path.join(tr.stats_dir, files[1]))
    assert sorted(vs) == [0, 24]
    # the saved best checkpoint loads back into a fresh agent
    agent2 = make_agent(trained=False)
    agent2.load_models(tr.best_validation_filepath[:-3])
