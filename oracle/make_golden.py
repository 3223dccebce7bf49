"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/* by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_import.py) in this container.

    python -m oracle.make_golden [stage ...]      stages: graphs env train forward learn rollouts trainloop hostutils replay

The fixtures are what travels to the GPU box (where /root/reference does not exist); every
fixture records the script stage and seeds that produced it.
"""
import os
import random
import sys
import time

import numpy as np
import torch

from oracle.ref_import import import_reference, load_reference_dataset

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
WEIGHTS = os.path.join(GOLD, 'parity_weights_GN')


def _adj_from_target(t):
    n = t['no_vertices']
    a = np.zeros((n, n), dtype=np.uint8)
    for (u, v), w in zip(t['edge_list'], t['edge_attr']):
        if w == -1:
            a[u, v] = 1
    return a


def stage_graphs(ref, ds):
    """The 100 validation graphs as dense adjacency + caller edge order + the stored DSATUR
    known answers + the neighbour-distance counts the first-vertex rule reads."""
    ns, adjs, edges, orders, used, names, ndc, fam = [], [], [], [], [], [], [], []
    for t, d in zip(ds.graphs, ds.dsatur_stats_by_graph):
        ns.append(t['no_vertices'])
        adjs.append(_adj_from_target(t).reshape(-1))
        edges.append(np.asarray(t['edge_list'], dtype=np.int16).reshape(-1))
        orders.append(np.asarray([int(v) for v in d['vertex_order']], dtype=np.int16))
        used.append(int(d['colours_used']))
        names.append(str(t['name']))
        ndc.append(np.asarray(t['neighbour_dist_counts'], dtype=np.int16).reshape(-1))
    np.savez_compressed(
        os.path.join(GOLD, 'validation_graphs.npz'),
        n=np.asarray(ns, dtype=np.int32), adj=np.concatenate(adjs), edge_list=np.concatenate(edges),
        dsatur_order=np.concatenate(orders), dsatur_order_len=np.asarray([len(o) for o in orders], dtype=np.int32),
        dsatur_colours_used=np.asarray(used, dtype=np.int32), names=np.asarray(names),
        neighbour_dist_counts=np.concatenate(ndc),
        provenance=np.asarray('oracle/make_golden.py stage_graphs from reference datasets/validation_dataset.pickle'))
    print('graphs: wrote', len(ns))


def stage_env(ref, ds):
    """Random action sequences through the reference env: colours / reward / done per step."""
    import gym
    rng = random.Random(1234)
    env = gym.make('graph-colouring-v0', dataset_graphs=ds.graphs)
    gids, acts, cols, rews, dones, lens = [], [], [], [], [], []
    for gi in range(0, 100, 3):
        t = ds.graphs[gi]
        _, cur = env.reset(target=t)
        done, steps = False, 0
        while not done:
            v = rng.choice(cur['unassigned_vertices'])
            cur, r, done, info = env.step(v)
            data, gf = env.GenerateData_v1(t, cur)
            assert gf == []
            acts.append(v)
            cols.append(data.x[:, 1].numpy().astype(np.int16))
            rews.append(int(r))
            dones.append(int(done))
            assert np.array_equal(np.asarray(cur['assigned_vertices_onehot']), (cols[-1] >= 0).astype(int))
            steps += 1
        gids.append(gi)
        lens.append(steps)
    np.savez_compressed(
        os.path.join(GOLD, 'env_traces.npz'), graph_ids=np.asarray(gids, dtype=np.int32),
        steps=np.asarray(lens, dtype=np.int32), actions=np.asarray(acts, dtype=np.int16),
        colours=np.concatenate(cols), rewards=np.asarray(rews, dtype=np.int16),
        dones=np.asarray(dones, dtype=np.int8),
        provenance=np.asarray('oracle/make_golden.py stage_env, random.Random(1234), reference GraphColouring.step'))
    print('env: wrote', len(gids), 'episodes', len(acts), 'steps')


def stage_train(ref, ds, episodes=160):
    """Briefly train the reference agent on CPU (dqn_main.py:158-199 loop, seeds 0, the
    validation graphs standing in for the absent training pickle) so that parity tests have
    weights with real Q gaps; the trained_policies/learned_parameters_GN blob is absent."""
    import gym
    random.seed(0)
    np.random.seed(0)
    torch.manual_seed(0)
    env = gym.make('graph-colouring-v0', dataset_graphs=ds.graphs)
    agent = ref.dqn_agent.DQNAgentGN(2, 1, 0)
    agent.no_episodes = 400        # faster epsilon decay than the 25 000-episode schedule
    t_step, t0 = 0, time.time()
    for ep in range(1, episodes + 1):
        _, target, cur = env.reset()
        data, gf = env.GenerateData_v1(target, cur)
        ep_step, done = 0, False
        while not done:
            if ep_step == 0:
                v = random.choice(cur['unassigned_vertices'])
            else:
                v = agent.choose_action(data, gf, cur['unassigned_vertices'])
            nxt, r, done, info = env.step(v)
            ndata, ngf = env.GenerateData_v1(target, nxt)
            if ep_step > 0:
                agent.remember(data, gf, np.array(nxt['assigned_vertices_onehot']), v, r, ndata, ngf, int(done))
            if t_step % agent.update_every == 0:
                agent.learn()
            t_step += 1
            data, gf, cur = ndata, ngf, nxt
            ep_step += 1
        agent.episode_no += 1
        if ep % 10 == 0:
            print('train ep', ep, 'colours', env.current_graph['no_colours_used'], 't=%.0fs' % (time.time() - t0), flush=True)
    path = agent.save_models(os.path.join(GOLD, 'parity_weights'))
    print('train: saved', path)


def _load_agent(ref, weights=True, root='parity_weights'):
    torch.manual_seed(0)
    agent = ref.dqn_agent.DQNAgentGN(2, 1, 0)
    if weights:
        agent.load_models(os.path.join(GOLD, root))
        agent.soft_update(agent.gn_behaviour, agent.gn_target, 1)
    return agent


def _partial_states(ref, ds, gids, seed):
    """Mid-episode observations: colour ~half the vertices in random order through the env."""
    import gym
    rng = random.Random(seed)
    env = gym.make('graph-colouring-v0', dataset_graphs=ds.graphs)
    out = []
    for gi in gids:
        t = ds.graphs[gi]
        _, cur = env.reset(target=t)
        k = max(1, t['no_vertices'] // 3)
        for _ in range(k):
            if not cur['unassigned_vertices']:
                break
            cur, r, done, info = env.step(rng.choice(cur['unassigned_vertices']))
        if not cur['unassigned_vertices']:       # finished early: restart, colour one vertex only
            _, cur = env.reset(target=t)
            cur, r, done, info = env.step(rng.choice(cur['unassigned_vertices']))
        data, _ = env.GenerateData_v1(t, cur)
        out.append((gi, data, list(cur['unassigned_vertices'])))
    return out


def stage_forward(ref, ds):
    """Reference GNNetwork outputs (q, h5 checksums) on mid-episode states, for the seed-0
    default init and for the briefly trained weights; plus init checksums per tensor."""
    res = {}
    for tag, use_w in (('init', False), ('trained', True)):
        agent = _load_agent(ref, use_w)
        net = agent.gn_behaviour
        sd = net.state_dict()
        res[tag + '_param_sum'] = np.asarray([float(v.double().sum()) for v in sd.values()])
        res[tag + '_param_abs'] = np.asarray([float(v.double().abs().sum()) for v in sd.values()])
        states = _partial_states(ref, ds, list(range(0, 100, 7)), seed=77)
        qs, hsum, gids, cols, acts = [], [], [], [], []
        for gi, data, un in states:
            with torch.no_grad():
                h, q = net(data, torch.zeros(1, 0), [0] * data.x.shape[0], [0] * data.edge_attr.shape[0])
            qs.append(q.view(-1).numpy().astype(np.float32))
            hsum.append(h.double().sum(1).numpy())
            gids.append(gi)
            cols.append(data.x[:, 1].numpy().astype(np.int16))
            acts.append(agent.choose_action(data, [], un, test_mode=True))
        res[tag + '_q'] = np.concatenate(qs)
        res[tag + '_h5_rowsum'] = np.concatenate(hsum)
        res[tag + '_actions'] = np.asarray(acts, dtype=np.int32)
        res['graph_ids'] = np.asarray(gids, dtype=np.int32)
        res['colours'] = np.concatenate(cols)
    res['keys'] = np.asarray(list(sd.keys()))
    res['provenance'] = np.asarray('oracle/make_golden.py stage_forward; torch.manual_seed(0) init; states seed 77')
    np.savez_compressed(os.path.join(GOLD, 'gn_forward.npz'), **res)
    print('forward: wrote', len(gids), 'states x 2 weight sets')


def stage_learn(ref, ds):
    """One reference DQNAgentGN.learn() on a fixed 8-transition minibatch: loss, every
    gradient tensor, post-Adam behaviour weights and post-soft-update target weights."""
    import gym
    agent = _load_agent(ref, True)
    agent.batch_size = 8
    # perturb the target net so behaviour != target
    with torch.no_grad():
        g = torch.Generator().manual_seed(5)
        for p in agent.gn_target.parameters():
            p.mul_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
    rng = random.Random(99)
    env = gym.make('graph-colouring-v0', dataset_graphs=ds.graphs)
    gids = [3, 10, 22, 37, 41, 58, 76, 93]
    trans = []
    for i, gi in enumerate(gids):
        t = ds.graphs[gi]
        _, cur = env.reset(target=t)
        k = 1 if i == 0 else rng.randrange(1, t['no_vertices'])
        done = False
        for _ in range(k):
            if len(cur['unassigned_vertices']) <= 1:
                break
            cur, r, done, info = env.step(rng.choice(cur['unassigned_vertices']))
        if done:   # ensure a non-terminal current state
            _, cur = env.reset(target=t)
            cur, r, done, info = env.step(rng.choice(cur['unassigned_vertices']))
        data, gf = env.GenerateData_v1(t, cur)
        v = rng.choice(cur['unassigned_vertices'])
        nxt, r, done, info = env.step(v)
        ndata, ngf = env.GenerateData_v1(t, nxt)
        trans.append((gi, data, np.array(nxt['assigned_vertices_onehot']), v, r, ndata, int(done)))
    # bypass random sampling: fill the buffer with exactly these transitions, in order
    mem = agent.memory
    for (gi, data, onehot, v, r, ndata, done) in trans:
        mem.store_transition(data, [], onehot, v, r, ndata, [], done)
    mem_sample = mem.sample_buffer
    mem.sample_buffer = lambda n: _ordered_sample(mem, len(trans))
    mem.current_mem_size_saved = mem.current_mem_size
    agent.memory.current_mem_size = 10 * agent.batch_size   # pass the gate at dqn_agent.py:149
    losses = {}
    import torch.nn.functional as F
    orig_mse = F.mse_loss

    def rec_mse(a, b, *args, **kw):
        out = orig_mse(a, b, *args, **kw)
        losses['loss'] = float(out)
        losses['q_expected'] = a.detach().view(-1).numpy().copy()
        losses['q_targets'] = b.detach().view(-1).numpy().copy()
        return out
    ref.dqn_agent.F.mse_loss = rec_mse
    before_t = [p.detach().clone() for p in agent.gn_target.parameters()]
    agent.learn()
    ref.dqn_agent.F.mse_loss = orig_mse
    mem.sample_buffer = mem_sample
    res = dict(
        graph_ids=np.asarray(gids, dtype=np.int32),
        cur_colours=np.concatenate([d.x[:, 1].numpy().astype(np.int16) for (_, d, *_r) in trans]),
        next_colours=np.concatenate([tr[5].x[:, 1].numpy().astype(np.int16) for tr in trans]),
        actions=np.asarray([tr[3] for tr in trans], dtype=np.int32),
        rewards=np.asarray([tr[4] for tr in trans], dtype=np.float32),
        dones=np.asarray([tr[6] for tr in trans], dtype=np.int32),
        loss=np.asarray(losses['loss']), q_expected=losses['q_expected'], q_targets=losses['q_targets'],
        target_before=np.concatenate([p.view(-1).numpy() for p in before_t]),
        grads=np.concatenate([p.grad.view(-1).numpy() for p in agent.gn_behaviour.parameters()]),
        behaviour_after=np.concatenate([p.detach().view(-1).numpy() for p in agent.gn_behaviour.parameters()]),
        target_after=np.concatenate([p.detach().view(-1).numpy() for p in agent.gn_target.parameters()]),
        provenance=np.asarray('oracle/make_golden.py stage_learn: reference DQNAgentGN.learn(), batch 8, target perturbed seed 5'))
    np.savez_compressed(os.path.join(GOLD, 'learn_step.npz'), **res)
    print('learn: loss', losses['loss'])


def _ordered_sample(mem, n):
    from torch_geometric.data import Batch
    idx = np.arange(n)
    return (Batch.from_data_list([mem.data_memory[i] for i in idx]), mem.current_global_features[idx],
            mem.assigned_vertices_memory[idx], mem.action_ind_memory[idx], mem.reward_memory[idx],
            Batch.from_data_list([mem.next_data_memory[i] for i in idx]), mem.next_global_features[idx],
            mem.done_memory[idx])


def stage_rollouts_trained25k(ref, ds):
    """The same rollouts with FULLY trained weights: tests/golden/trained25k_GN is the behaviour network after the
    reference's whole 25 000-episode schedule run by this repository's device training loop on a B200
    (scripts/train_full_schedule.py, profiles/r02_train_full_schedule.json) -- the closest available stand-in for the
    reference's missing trained_policies/learned_parameters_GN.  Here the UNMODIFIED reference plays them on its 100
    validation graphs on the CPU; tests/test_trained_policy_gpu.py replays them on the device."""
    stage_rollouts(ref, ds, root='trained25k', out='rollouts_trained25k.npz')


def stage_rollouts(ref, ds, root='parity_weights', out='rollouts.npz'):
    """Reference greedy rollouts (test_policy_functions.test_using_learned_policy) with the
    trained weights on all 100 validation graphs: vertex sequences, colours used, and the
    top-2 Q gap at every NN-chosen step."""
    agent = _load_agent(ref, True, root)
    seqs, gaps = [], []
    cur = {'seq': None}
    EnvCls = ref.env.GraphColouring
    orig_step = EnvCls.step
    orig_choose = agent.choose_action

    def rec_step(self, v):
        cur['seq'].append(int(v))
        return orig_step(self, v)

    def rec_choose(data, gf, un, epsilon=None, test_mode=False):
        with torch.no_grad():
            _, q = agent.gn_behaviour(data, torch.zeros(1, 0), [0] * data.x.shape[0], [0])
        qs = np.sort(q.view(-1).numpy()[un])
        cur['gaps'].append(float(qs[-1] - qs[-2]) if len(qs) > 1 else np.inf)
        return orig_choose(data, gf, un, epsilon=epsilon, test_mode=test_mode)
    EnvCls.step = rec_step
    agent.choose_action = rec_choose
    used = []
    t0 = time.time()
    for t in ds.graphs:
        cur['seq'], cur['gaps'] = [], []
        stats, by = ref.test_policy_functions.test_using_learned_policy([t], agent, False)
        used.append(int(by[0]['min_colours_used']))
        seqs.append(np.asarray(cur['seq'], dtype=np.int16))
        gaps.append(np.asarray(cur['gaps'], dtype=np.float32))
    EnvCls.step = orig_step
    np.savez_compressed(
        os.path.join(GOLD, out), colours_used=np.asarray(used, dtype=np.int32),
        seq=np.concatenate(seqs), seq_len=np.asarray([len(s) for s in seqs], dtype=np.int32),
        gaps=np.concatenate(gaps), gaps_len=np.asarray([len(g) for g in gaps], dtype=np.int32),
        provenance=np.asarray('oracle/make_golden.py stage_rollouts: reference test_using_learned_policy, %s_GN' % root))
    print('rollouts: mean colours %.3f in %.0fs' % (np.mean(used), time.time() - t0))


def stage_trainloop(ref, ds, episodes=24):
    """The reference training loop (dqn_main.py:158-199) for a few episodes from the parity weights:
    per-episode graph index, action sequence (with a flag for NN-chosen actions and their top-2 Q
    gap), colours used, the loss of every learn() and weight checksums at the end.  batch_size 8
    (so the 10 x batch gate opens after ~4 episodes) and a 60-episode epsilon schedule."""
    import gym
    import torch.nn.functional as F
    agent = _load_agent(ref, True)
    agent.batch_size = 8
    agent.no_episodes = 60
    random.seed(3)
    np.random.seed(3)
    torch.manual_seed(3)
    env = gym.make('graph-colouring-v0', dataset_graphs=ds.graphs)
    losses = []
    orig_mse = F.mse_loss

    def rec_mse(a, b, *args, **kw):
        out = orig_mse(a, b, *args, **kw)
        losses.append(float(out))
        return out
    ref.dqn_agent.F.mse_loss = rec_mse
    seen = {}

    def hook(mod, inp, out):
        seen['q'] = out[1].detach().view(-1).numpy().copy()
    h = agent.gn_behaviour.register_forward_hook(hook)
    gids, acts, flags, gaps, lens, used, scores = [], [], [], [], [], [], []
    t_step, t0 = 0, time.time()
    for ep in range(1, episodes + 1):
        gi, target, cur = env.reset()
        data, gf = env.GenerateData_v1(target, cur)
        ep_step, done, score = 0, False, 0
        while not done:
            if ep_step == 0:
                v = random.choice(cur['unassigned_vertices'])
                flags.append(0)
                gaps.append(np.inf)
            else:
                seen.pop('q', None)
                v = agent.choose_action(data, gf, cur['unassigned_vertices'])
                if 'q' in seen:
                    qs = np.sort(seen['q'][cur['unassigned_vertices']])
                    flags.append(1)
                    gaps.append(float(qs[-1] - qs[-2]) if len(qs) > 1 else np.inf)
                else:
                    flags.append(0)
                    gaps.append(np.inf)
            acts.append(int(v))
            nxt, r, done, info = env.step(v)
            ndata, ngf = env.GenerateData_v1(target, nxt)
            if ep_step > 0:
                agent.remember(data, gf, np.array(nxt['assigned_vertices_onehot']), v, r, ndata, ngf, int(done))
            if t_step % agent.update_every == 0:
                agent.learn()
            t_step += 1
            data, gf, cur = ndata, ngf, nxt
            score += r
            ep_step += 1
        agent.episode_no += 1
        gids.append(gi)
        lens.append(ep_step)
        used.append(env.current_graph['no_colours_used'])
        scores.append(score)
    h.remove()
    ref.dqn_agent.F.mse_loss = orig_mse
    sd_b, sd_t = agent.gn_behaviour.state_dict(), agent.gn_target.state_dict()
    np.savez_compressed(
        os.path.join(GOLD, 'train_loop.npz'), graph_ids=np.asarray(gids, dtype=np.int32),
        ep_len=np.asarray(lens, dtype=np.int32), actions=np.asarray(acts, dtype=np.int32),
        nn_chosen=np.asarray(flags, dtype=np.int8), gaps=np.asarray(gaps, dtype=np.float32),
        colours_used=np.asarray(used, dtype=np.int32), scores=np.asarray(scores, dtype=np.int32),
        losses=np.asarray(losses, dtype=np.float64), buffer_size=np.asarray(agent.memory.current_mem_size),
        behaviour_sum=np.asarray([float(v.double().sum()) for v in sd_b.values()]),
        behaviour_abs=np.asarray([float(v.double().abs().sum()) for v in sd_b.values()]),
        target_sum=np.asarray([float(v.double().sum()) for v in sd_t.values()]),
        target_abs=np.asarray([float(v.double().abs().sum()) for v in sd_t.values()]),
        provenance=np.asarray('oracle/make_golden.py stage_trainloop: reference dqn_main.py loop, parity_weights_GN, '
                              'batch_size 8, no_episodes 60, seeds 3, %d episodes' % episodes))
    print('trainloop: %d episodes, %d steps, %d learns in %.0fs' % (episodes, len(acts), len(losses), time.time() - t0))


def stage_hostutils(ref, ds):
    """Reference host helpers on small inputs: DIMACS reader, queen graphs, networkx conversion and the
    three stat-file writers (utils.py:58-125,243-323, construct_queen_graph.py)."""
    import json
    import tempfile
    from collections import namedtuple
    sys.path.insert(0, os.environ.get('GCRL_REFERENCE_ROOT', '/root/reference'))
    import construct_queen_graph as cq
    U = ref.utils
    col = ('c tiny DIMACS file\nc with a self loop, a duplicate and both directions\np edge 7 9\n'
           'e 1 2\ne 2 1\ne 1 7\ne 3 4\ne 4 5\ne 5 5\ne 6 3\ne 2 6\ne 1 2\n')
    out = {'dimacs_text': col}
    with tempfile.TemporaryDirectory() as d:
        fp = os.path.join(d, 'tiny.col')
        with open(fp, 'w') as f:
            f.write(col)
        n, el, ea, ind, idx = U.import_dimacs_graph(fp)
        out['dimacs'] = dict(n=n, edge_list=[list(e) for e in el], edge_attr=list(ea),
                             ind=[[u, v, k] for u, dd in ind.items() for v, k in dd.items()], indices=sorted(idx))
        g = U.import_dimacs_graph_as_networkx(fp)
        out['dimacs_nx_edges'] = sorted([sorted(e) for e in g.edges()])
        queens = []
        for seed, (n, k) in enumerate([(12, 3), (12, 4), (20, 4), (25, 5), (30, 6), (16, 2)]):
            random.seed(seed)
            el, ea, ind, idx = cq.generate_queen_graph(n, k)
            queens.append(dict(seed=seed, n=n, k=k, edge_list=[list(e) for e in el], edge_attr=list(ea), indices=list(idx),
                               after=random.random()))
        out['queens'] = queens
        import construct_leighton_graph as cl
        leighton = []
        for seed, (n, k) in enumerate([(10, 5), (12, 3), (20, 4), (30, 5), (45, 9)]):
            random.seed(100 + seed)
            el, ea, ind, idx = cl.generate_leighton_graph(n, k)
            leighton.append(dict(seed=100 + seed, n=n, k=k, edge_attr=list(ea), indices=list(idx), after=random.random()))
        random.seed(7)
        m, c, a, b = cl.generate_parameters(24, 6)
        out['leighton'] = leighton
        out['leighton_params'] = dict(seed=7, n=24, k=6, m=m, c=c, a=a, b=b)
        verts, edges = ['a', 'b', 'c', 'd', 'e'], [('a', 'c'), ('e', 'b'), ('c', 'd'), ('d', 'a')]
        el, ea, ind, idx = U.convert_networkx_to_graph(verts, edges)
        out['convert'] = dict(vertices=verts, edges=[list(e) for e in edges], edge_list=[list(e) for e in el],
                              edge_attr=list(ea), indices=list(idx),
                              ind=[[u, v, k] for u, dd in ind.items() for v, k in dd.items()])
        T = namedtuple("Stats", ["saved_episodes", "episode_rewards", "colours_used"])
        V = namedtuple("Stats", ["episode_rewards", "colours_used"])
        tr = T([1, 2, 3], [-7, -5, -6], [7, 5, 6])
        va = {0: V([-7.0, -5.0], [7.0, 5.0]), 500: V([-6.0, -4.0], [6.0, 4.0])}
        be = V([-8.25, -6.5], [8.25, 6.5])
        U.save_training_stats_to_file(tr, os.path.join(d, 't'))
        U.save_validation_stats_to_file(va, os.path.join(d, 'v'))
        U.save_benchmark_stats_to_file(be, os.path.join(d, 'b'))
        out['stat_files'] = {k: open(os.path.join(d, k)).read() for k in 'tvb'}
    out['provenance'] = 'oracle/make_golden.py stage_hostutils (reference utils.py / construct_queen_graph.py)'
    with open(os.path.join(GOLD, 'host_utils.json'), 'w') as f:
        json.dump(out, f)
    print('hostutils: wrote', os.path.getsize(os.path.join(GOLD, 'host_utils.json')), 'bytes')


def stage_replay(ref, ds):
    """The reference ReplayBuffer (dqn_agent.py:8-65) with 11 slots fed 29 env transitions (the ring wraps twice), then
    three seeded sample_buffer calls: what each minibatch returned (colours before / after, actions, rewards, dones,
    assigned one-hots, batched edge_attr checksum).  Pins oracle/replay_oracle.py."""
    import gym
    rng = random.Random(17)
    env = gym.make('graph-colouring-v0', dataset_graphs=ds.graphs)
    mem = ref.dqn_agent.ReplayBuffer(11, 0)
    gids, cur_c, nxt_c, acts, rews, dones = [], [], [], [], [], []
    gi_list = [2, 31, 64, 77]
    k = 0
    while k < 29:
        gi = gi_list[k % len(gi_list)]
        t = ds.graphs[gi]
        _, cur = env.reset(target=t)
        data, gf = env.GenerateData_v1(t, cur)
        done = False
        while not done and k < 29:
            v = rng.choice(cur['unassigned_vertices'])
            nxt, r, done, info = env.step(v)
            ndata, ngf = env.GenerateData_v1(t, nxt)
            mem.store_transition(data, gf, np.array(nxt['assigned_vertices_onehot']), v, r, ndata, ngf, int(done))
            gids.append(gi)
            cur_c.append(data.x[:, 1].numpy().astype(np.int16))
            nxt_c.append(ndata.x[:, 1].numpy().astype(np.int16))
            acts.append(int(v)); rews.append(float(r)); dones.append(int(done))
            data, gf, cur = ndata, ngf, nxt
            k += 1
    out = dict(capacity=np.asarray(11), graph_ids=np.asarray(gids, dtype=np.int32), cur=np.concatenate(cur_c),
               nxt=np.concatenate(nxt_c), actions=np.asarray(acts, dtype=np.int32), rewards=np.asarray(rews, dtype=np.float32),
               dones=np.asarray(dones, dtype=np.int32), count=np.asarray(mem.current_mem_size))
    np.random.seed(23)
    for b in range(3):
        d, gfs, assigned, a, r, nd, ngfs, dn = mem.sample_buffer(6)
        out['s%d_x' % b] = d.x[:, 1].numpy().astype(np.int16)
        out['s%d_nx' % b] = nd.x[:, 1].numpy().astype(np.int16)
        out['s%d_nodes' % b] = np.asarray([int(n) for n in np.bincount(d.batch.numpy())], dtype=np.int32)
        out['s%d_actions' % b] = a.reshape(-1).astype(np.int32)
        out['s%d_rewards' % b] = r.astype(np.float32)
        out['s%d_dones' % b] = dn.astype(np.int32)
        out['s%d_assigned' % b] = np.concatenate([np.asarray(x) for x in assigned]).astype(np.int8)
        out['s%d_present_edges' % b] = np.asarray(int((d.edge_attr == -1).sum()))
    out['provenance'] = np.asarray('oracle/make_golden.py stage_replay: reference ReplayBuffer(11, 0), 29 transitions, np.random.seed(23), 3 x sample_buffer(6)')
    np.savez_compressed(os.path.join(GOLD, 'replay_buffer.npz'), **out)
    print('replay: stored', mem.current_mem_size)


def main(argv):
    stages = argv or ['graphs', 'env', 'train', 'forward', 'learn', 'rollouts']
    os.makedirs(GOLD, exist_ok=True)
    ref = import_reference()
    ds = load_reference_dataset()
    for s in stages:
        globals()['stage_' + s](ref, ds)


if __name__ == '__main__':
    main(sys.argv[1:])
