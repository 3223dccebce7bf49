"""The graph-colouring environment on the device.

`GraphColouring` is a drop-in for gym_graph_colouring.envs.GraphColouring
(graph_colouring_env.py:16-191): `reset(target=None)`, `step(vertex)`,
`GenerateData_v1(target, current_graph)`, same return shapes.  The state lives in device
arrays advanced by `gcrl_env_step`; the `current_graph` dict handed back to the caller holds
the keys the reference's callers read (dqn_main.py:165-199, test_policy_functions.py:33-73).

`BatchedColouringEnv` is the same environment for G graphs at once with no host round trip
(one CTA per graph per step); greedy rollouts over thousands of graphs use it.
"""
import random

import numpy as np
import torch

from . import ops
from .architecture import _device
from .data import Data

MAX_RADIUS_FOR_NEIGHBOURS = 3      # utils.py:19
ENV_MAX_N = 4096                   # vertices per graph the env kernels support (csrc/env.cu ENV_MAX_N; colours are int16 in the replay ring)


def check_graph_sizes(ns):
    """The env kernels keep a graph's colours in shared memory: a graph over ENV_MAX_N vertices would be skipped
    silently (state never advances, `done` never set), so it is rejected here."""
    ns = np.asarray(ns)
    if ns.size and int(ns.max()) > ENV_MAX_N:
        raise ValueError('graph with %d vertices: the device environment supports at most %d' % (int(ns.max()), ENV_MAX_N))
    if ns.size and int(ns.min()) < 1:
        raise ValueError('graph without vertices')


class ColouringData(Data):
    """Observation of one graph: x = [vertex name, colour], the static complete edge list with
    edge_attr -1/0 (graph_colouring_env.py:124-141) plus the dense adjacency `adj` [n,n] uint8
    the device batching path uses."""
    pass


def adjacency_from_target(target):
    """Dense 0/1 adjacency from a reference target dict (edge_attr == -1 marks a present edge,
    utils.py:296-323)."""
    if 'adj' in target:
        return np.ascontiguousarray(target['adj'], dtype=np.uint8)
    n = int(target['no_vertices'])
    el = np.asarray(target['edge_list'], dtype=np.int64).reshape(-1, 2)
    ea = np.asarray(target['edge_attr']).reshape(-1)
    adj = np.zeros((n, n), dtype=np.uint8)
    present = el[ea == -1]
    adj[present[:, 0], present[:, 1]] = 1
    return adj


def complete_edge_list(n):
    """(0,1),(1,0),(0,2),(2,0),... -- the canonical order of utils.py:304-310, vectorised."""
    a, b = np.triu_indices(n, 1)
    out = np.empty((a.size * 2, 2), dtype=np.int64)
    out[0::2, 0], out[0::2, 1] = a, b
    out[1::2, 0], out[1::2, 1] = b, a
    return out


def make_target(adj, name='graph', device=None):
    """Target dict with the keys the hot path reads: no_vertices, vertex_names,
    edge_list, edge_attr, edge_ind_dict, present_edges, neighbour_dist_counts, name (+ adj).
    The <=3-hop neighbour counts come from gcrl_env_first_vertex on the device."""
    adj = np.ascontiguousarray(adj, dtype=np.uint8)
    n = adj.shape[0]
    el = complete_edge_list(n)
    ea = -adj[el[:, 0], el[:, 1]].astype(np.int64)
    dev = _device() if device is None else torch.device(device)
    a = torch.from_numpy(adj.reshape(-1)).to(dev)
    node_off = torch.tensor([0, n], dtype=torch.int32, device=dev)
    adj_off = torch.zeros(1, dtype=torch.int64, device=dev)
    _, counts = ops.env_first_vertex(a, adj_off, node_off, n, want_counts=True)
    edge_ind = {}
    for k, (u, v) in enumerate(el.tolist()):
        edge_ind.setdefault(u, {})[v] = k
    return {
        'name': name, 'no_vertices': n, 'vertex_names': list(range(n)),
        'edge_list': [tuple(e) for e in el.tolist()], 'edge_attr': ea.tolist(), 'edge_ind_dict': edge_ind,
        'present_edges': [tuple(e) for e in el[ea == -1].tolist()],
        'neighbour_dist_counts': counts.cpu().numpy().astype(np.float64), 'adj': adj,
    }


class _GraphOnDevice(object):
    """Static per-graph device data, built once per target."""

    def __init__(self, target, device):
        self.target = target
        adj = adjacency_from_target(target)
        self.n = n = adj.shape[0]
        check_graph_sizes([n])
        self.adj_np = adj
        self.adj = torch.from_numpy(adj.reshape(-1)).to(device)
        self.adj_off = torch.zeros(1, dtype=torch.int64, device=device)
        self.node_off = torch.tensor([0, n], dtype=torch.int32, device=device)
        el = np.asarray(target['edge_list'], dtype=np.int64).reshape(-1, 2)
        self.edge_index = torch.from_numpy(np.ascontiguousarray(el.T)).to(device)
        self.edge_attr = torch.tensor(np.asarray(target['edge_attr'], dtype=np.float32).reshape(-1, 1), device=device)
        self.names = torch.arange(n, dtype=torch.float32, device=device)


class GraphColouring(object):
    """Single-graph environment with the reference's interface."""
    metadata = {'render.modes': ['human']}

    def __init__(self, dataset_graphs):
        self.graphs = dataset_graphs
        self.device = _device()
        self._cache = {}
        self.target = None
        self.current_graph = None

    def _static(self, target):
        g = self._cache.get(id(target))
        if g is None or g.target is not target:
            g = self._cache[id(target)] = _GraphOnDevice(target, self.device)
        return g

    def reset(self, target=None):
        self.episode_step = 0
        if target is None:
            target_ind = random.choice(range(len(self.graphs)))      # graph_colouring_env.py:48-50
            self.target = self.graphs[target_ind]
        else:
            self.target = target
        g = self._g = self._static(self.target)
        self.colour = torch.empty(g.n, dtype=torch.int32, device=self.device)
        self.max_colour = torch.empty(1, dtype=torch.int32, device=self.device)
        self.x = torch.empty(g.n, 2, dtype=torch.float32, device=self.device)
        self._action = torch.empty(1, dtype=torch.int32, device=self.device)
        ops.env_reset(self.colour, self.max_colour, self.x, g.node_off)
        self.current_graph = self._host_state(-np.ones(g.n, dtype=np.int64))
        if target is None:
            return target_ind, self.target, self.current_graph
        return self.target, self.current_graph

    def _host_state(self, col):
        un = [int(v) for v in np.nonzero(col < 0)[0]]
        used = int(col.max()) + 1 if col.size else 0
        return {
            'vertex_values': {v: int(c) for v, c in enumerate(col.tolist())},
            'assigned_vertices_onehot': (col >= 0).astype(np.int64).tolist(),
            'unassigned_vertices': un,
            'no_coloured_vertices': int(col.size - len(un)),
            'no_uncoloured_vertices': len(un),
            'no_colours_used': max(used, 0),
            'max_colour_ind_used': max(used, 0) - 1,
            'colour': col,
        }

    def step(self, action_vertex):
        g = self._g
        self.episode_step += 1
        if not (isinstance(action_vertex, (int, np.integer)) and 0 <= int(action_vertex) < g.n):
            raise Exception(str(action_vertex) + ' is not a vertex name.')      # :72-73
        prev_used = self.current_graph['no_colours_used']
        self._action.fill_(int(action_vertex))
        ops.env_step(g.adj, g.adj_off, g.node_off, self._action, self.colour, self.max_colour, None, None, self.x)
        col = self.colour.cpu().numpy().astype(np.int64)                        # the one sync per step
        self.current_graph = self._host_state(col)
        self.done = (self.current_graph['no_uncoloured_vertices'] == 0)         # :82
        rewards = self.GenerateRewards(self.current_graph['no_colours_used'], prev_used)
        info = {'rewards': rewards, 'episode_ended': self.done}
        return (self.current_graph, sum(rewards.values()), self.done, info)

    def GenerateRewards(self, no_colours_used, prev_no_colours_used):
        return {'new_colours_used': -1 * (no_colours_used - prev_no_colours_used)}   # :102-109

    def GenerateData_v1(self, target, current_graph):
        """-> (Data, []) (graph_colouring_env.py:112-141).  The Data lives on the device; its
        edge_index / edge_attr are the graph's static tensors (batched once, reused every step)."""
        g = self._static(target)
        col = current_graph['colour'] if 'colour' in current_graph else np.asarray(
            [int(current_graph['vertex_values'][v]) for v in target['vertex_names']])
        x = torch.stack([g.names, torch.from_numpy(np.asarray(col, dtype=np.float32)).to(self.device)], 1)
        data = ColouringData(x=x, edge_index=g.edge_index, edge_attr=g.edge_attr)
        data.adj = g.adj
        # the list this observation's x[:,1] < 0 mask was derived from (choose_action's device-mask fast path)
        data._unassigned = current_graph.get('unassigned_vertices') if 'colour' in current_graph else None
        return data, []

    def render(self, mode='human', close=False):
        print('target:', self.target.get('name'))
        print('current vertex values:', self.current_graph['vertex_values'])


def make(env_id, **kwargs):
    """gym.make stand-in: 'graph-colouring-v0' (gym_graph_colouring/__init__.py:1-5) and the
    unversioned id the rollout harness asks for (test_policy_functions.py:13)."""
    if env_id not in ('graph-colouring-v0', 'graph-colouring'):
        raise ValueError('unknown environment id %r' % (env_id,))
    return GraphColouring(**kwargs)


class BatchedColouringEnv(object):
    """G independent colouring episodes advanced together on the device."""

    def __init__(self, adjs, device=None):
        """adjs: list of [n,n] 0/1 arrays (or reference target dicts)."""
        dev = _device() if device is None else torch.device(device)
        mats = [adjacency_from_target(a) if isinstance(a, dict) else np.ascontiguousarray(a, dtype=np.uint8) for a in adjs]
        ns = np.asarray([m.shape[0] for m in mats], dtype=np.int64)
        adj = torch.from_numpy(np.concatenate([m.reshape(-1) for m in mats])).to(dev)
        adj_off = torch.from_numpy(np.concatenate([[0], np.cumsum(ns * ns)[:-1]]).astype(np.int64)).to(dev)
        self._setup(ns, adj, adj_off, dev)

    @classmethod
    def from_flat(cls, ns, flat_adj, device=None):
        """ns int64 [G] vertex counts, flat_adj uint8 [sum n^2] (ideally pinned host memory): one async H2D copy,
        no per-graph Python objects."""
        dev = _device() if device is None else torch.device(device)
        ns = np.asarray(ns, dtype=np.int64)
        self = cls.__new__(cls)
        adj = flat_adj.to(dev, non_blocking=True)
        adj_off = torch.from_numpy(np.concatenate([[0], np.cumsum(ns * ns)[:-1]]).astype(np.int64)).to(dev)
        self._setup(ns, adj, adj_off, dev)
        return self

    @classmethod
    def from_pool(cls, pool, graph_ids):
        """Environment batch over graphs that already live on the device (replay.GraphPool): the
        adjacency bytes are shared with the pool, only the per-graph offsets are built."""
        self = cls.__new__(cls)
        ids = np.asarray(graph_ids, dtype=np.int64)
        adj_off = torch.from_numpy(pool.adj_off_h[ids].copy()).to(pool.device)
        self._setup(pool.ns[ids], pool.adj, adj_off, pool.device)
        self.graph_ids = ids
        return self

    def _setup(self, ns, adj, adj_off, dev):
        check_graph_sizes(ns)
        self.device = dev
        self.n_graphs = int(ns.size)
        self.ns = ns
        self.n_vertices = int(ns.sum())
        self.n_edges = int((ns * (ns - 1)).sum())
        self.adj, self.adj_off = adj, adj_off
        self.node_off = torch.from_numpy(np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)).to(dev)
        G, N = self.n_graphs, self.n_vertices
        self.colour = torch.empty(N, dtype=torch.int32, device=dev)
        self.max_colour = torch.empty(G, dtype=torch.int32, device=dev)
        self.reward = torch.zeros(G, dtype=torch.float32, device=dev)
        self.done = torch.zeros(G, dtype=torch.int32, device=dev)
        self.x = torch.empty(N, 2, dtype=torch.float32, device=dev)
        self._packed = None
        self.reset()

    @property
    def packed(self):
        """(Packed, eattr) of the whole batch, built once by gcrl_pack_dense."""
        if self._packed is None:
            self._packed = ops.pack_dense(self.adj, self.adj_off, self.node_off, self.n_vertices, self.n_edges)
        return self._packed

    def reset(self):
        ops.env_reset(self.colour, self.max_colour, self.x, self.node_off)
        self.done.zero_()
        self.reward.zero_()

    def first_vertices(self):
        """test_policy_functions.py:46-47 for every graph -> int32 [G]."""
        return ops.env_first_vertex(self.adj, self.adj_off, self.node_off, self.n_vertices)

    def step(self, actions):
        """actions int32 [G] (local vertex ids; -1 = leave the graph alone)."""
        ops.env_step(self.adj, self.adj_off, self.node_off, actions, self.colour, self.max_colour, self.reward,
                     self.done, self.x)

    def step_greedy(self, q):
        """Argmax-Q over the uncoloured vertices of every graph fused with the step -> the actions taken, int32 [G]."""
        return ops.env_step_greedy(self.adj, self.adj_off, self.node_off, q, self.colour, self.max_colour, self.reward,
                                   self.done, self.x)

    def colours_used(self):
        return self.max_colour + 1
