"""CPU tests of host-side pieces of the drop-in that need no device: the Data / Batch containers (PyG surface the
reference touches), the reference target-dict <-> adjacency helpers, epsilon schedule and learn cadence arithmetic of the
device training loop."""
import numpy as np
import pytest
import torch

import gcrl_testutil as util

GRAPHS = util.load_validation_graphs()


def test_data_and_batch_follow_the_pyg_surface_the_reference_uses():
    from graph_colouring_with_rl_b200.data import Batch, Data
    g0, g1 = GRAPHS[0], GRAPHS[3]
    ds = []
    for g in (g0, g1):
        el = torch.from_numpy(g['edge_list'].T.copy())
        ea = torch.from_numpy(-g['adj'][g['edge_list'][:, 0], g['edge_list'][:, 1]].astype(np.float32)).view(-1, 1)
        x = torch.stack([torch.arange(g['n']).float(), -torch.ones(g['n'])], 1)
        ds.append(Data(x=x, edge_index=el, edge_attr=ea))
    b = Batch.from_data_list(ds)                                    # dqn_agent.py:55-56
    n0, e0 = g0['n'], g0['n'] * (g0['n'] - 1)
    assert b.x.shape == (g0['n'] + g1['n'], 2) and b.num_graphs == 2
    assert torch.equal(b.edge_index[:, :e0], ds[0].edge_index)
    assert torch.equal(b.edge_index[:, e0:], ds[1].edge_index + n0)                 # offset by the cumulative node count
    assert torch.equal(b.edge_attr, torch.cat([d.edge_attr for d in ds], 0))
    assert b.batch.tolist() == [0] * n0 + [1] * g1['n'] and b.ptr.tolist() == [0, n0, n0 + g1['n']]
    c = ds[0].clone()
    c.x[0, 1] = 5
    assert ds[0].x[0, 1] == -1 and ds[0].num_nodes == n0 and ds[0].num_edges == e0
    assert ds[0].to('cpu') is ds[0] and ds[0].cpu() is ds[0]                        # in place, returns self (PyG semantics)
    assert 'edge_index=[2, %d]' % e0 in repr(ds[0])


def test_target_dict_helpers_round_trip_the_reference_encoding():
    from graph_colouring_with_rl_b200.envs import adjacency_from_target, complete_edge_list
    from graph_colouring_with_rl_b200.utils import complete_edge_arrays, encode_adjacency
    for g in (GRAPHS[1], GRAPHS[40], GRAPHS[99]):
        el, ea, ind = encode_adjacency(g['adj'])
        t = {'no_vertices': g['n'], 'edge_list': el, 'edge_attr': ea}
        assert np.array_equal(adjacency_from_target(t), g['adj'])                   # edge_attr == -1 marks an edge (utils.py:296-323)
        assert np.array_equal(adjacency_from_target({'adj': g['adj']}), g['adj'])
        assert np.array_equal(complete_edge_list(g['n']), complete_edge_arrays(g['n']))
        # the fixture's own edge order (taken from the reference pickle) decodes to the same adjacency
        t2 = {'no_vertices': g['n'], 'edge_list': [tuple(e) for e in g['edge_list'].tolist()],
              'edge_attr': (-g['adj'][g['edge_list'][:, 0], g['edge_list'][:, 1]].astype(int)).tolist()}
        assert np.array_equal(adjacency_from_target(t2), g['adj'])


def test_epsilon_schedule_and_learn_cadence():
    from graph_colouring_with_rl_b200.dqn_main import epsilon_of

    class A(object):
        eps_start, eps_end, no_episodes, update_every = .9, .01, 25000, 5
    a = A()
    assert epsilon_of(a, 0) == pytest.approx(0.9) and epsilon_of(a, 25000) == pytest.approx(0.01)
    decay = (a.eps_end / a.eps_start) ** (1 / a.no_episodes)                         # dqn_agent.py:112-117
    assert epsilon_of(a, 1234) == pytest.approx(a.eps_start * decay ** 1234)
    # learn() once per update_every environment steps (dqn_main.py:186-188): batched steps of k environments run as many
    # learn steps as the reference's per-step test `t_step % update_every == 0` would have fired
    t, fired = 0, 0
    rng = np.random.default_rng(0)
    for _ in range(200):
        k = int(rng.integers(1, 70))
        first = -(-t // a.update_every) * a.update_every
        fired += len(range(first, t + k, a.update_every))
        t += k
    assert fired == len([s for s in range(t) if s % a.update_every == 0])
