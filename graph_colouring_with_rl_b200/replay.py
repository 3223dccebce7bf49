"""Device-resident replay memory (reference: dqn_agent.py:8-65 ReplayBuffer).

The reference stores two PyG `Data` objects per transition on the host and collates 2 x 64 of
them on the CPU at every learn().  Here the graphs live once in a `GraphPool` on the device and a
transition is (graph id, colours before, colours after, action, reward, done) in a ring of device
arrays written by `gcrl_replay_store`; `gather()` materialises a minibatch with one
`gcrl_replay_gather` launch as a `trainer.TransitionBatch`, i.e. directly in the layout the
learn step consumes.  Ring index and sampling follow the reference: slot = count % capacity
(:33), `np.random.choice(min(count, capacity), batch)` with replacement (:47-48).
"""
import numpy as np
import torch

from . import ops
from .architecture import _device
from .envs import adjacency_from_target, check_graph_sizes
from .trainer import TransitionBatch


class GraphPool(object):
    """The graphs of a dataset as dense adjacency bytes on the device."""

    def __init__(self, graphs, device=None):
        """graphs: list of [n,n] 0/1 arrays or reference target dicts."""
        self.device = dev = _device() if device is None else torch.device(device)
        mats = [adjacency_from_target(g) if isinstance(g, dict) else np.ascontiguousarray(g, dtype=np.uint8)
                for g in graphs]
        self.graphs = graphs
        self.ns = np.asarray([m.shape[0] for m in mats], dtype=np.int64)
        check_graph_sizes(self.ns)
        self.n_max = int(self.ns.max())
        self.adj_off_h = np.concatenate([[0], np.cumsum(self.ns * self.ns)[:-1]]).astype(np.int64)
        self.adj = torch.from_numpy(np.concatenate([m.reshape(-1) for m in mats])).to(dev)
        self.adj_off = torch.from_numpy(self.adj_off_h).to(dev)
        self.n = torch.from_numpy(self.ns.astype(np.int32)).to(dev)

    def __len__(self):
        return int(self.ns.size)


class DeviceReplayBuffer(object):
    def __init__(self, max_size, pool, stride=None):
        self.max_mem_size = int(max_size)
        self.current_mem_size = 0
        self.pool = pool
        dev = self.device = pool.device
        self.stride = int(stride or pool.n_max)
        if self.stride < pool.n_max:
            raise ValueError('stride %d is smaller than the largest graph (%d vertices)' % (self.stride, pool.n_max))
        C = self.max_mem_size
        self.ring = (torch.full((C,), -1, dtype=torch.int32, device=dev),
                     torch.full((C, self.stride), -1, dtype=torch.int16, device=dev),
                     torch.full((C, self.stride), -1, dtype=torch.int16, device=dev),
                     torch.zeros(C, dtype=torch.int32, device=dev),
                     torch.zeros(C, dtype=torch.float32, device=dev),
                     torch.zeros(C, dtype=torch.int32, device=dev))
        self.slot_graph_h = np.full(C, -1, dtype=np.int64)      # host mirror: which graph each slot holds
        self._err = torch.zeros(1, dtype=torch.int32, device=dev)

    def bytes(self):
        return int(sum(t.numel() * t.element_size() for t in self.ring))

    def store_from_env(self, env, graph_ids, cur_colour, action, mask=None):
        """Store the transition every environment of the batched `env` just made (cur_colour: its
        colours before the step; env.colour / reward / done: after).  mask: host bool [n_envs],
        environments to record (None = all).  Returns the slots written (host int array)."""
        G = env.n_graphs
        mask = np.ones(G, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        k = int(mask.sum())
        slot_h = np.full(G, -1, dtype=np.int32)
        slots = (self.current_mem_size + np.arange(k)) % self.max_mem_size          # dqn_agent.py:33
        if k > self.max_mem_size:
            raise ValueError('more transitions in one store than ring slots')
        slot_h[mask] = slots
        gids = np.asarray(graph_ids, dtype=np.int64)
        self.slot_graph_h[slots] = gids[mask]
        self.current_mem_size += k
        if k == 0:
            return slots
        dev = self.device
        ops.replay_store(torch.from_numpy(slot_h).to(dev), torch.from_numpy(gids.astype(np.int32)).to(dev),
                         env.node_off, cur_colour, env.colour, action, env.reward, env.done, self.ring)
        return slots

    def sample_indices(self, sample_size):
        max_mem = min(self.current_mem_size, self.max_mem_size)
        return np.random.choice(max_mem, sample_size)                               # :47-48

    def gather(self, batch):
        """Ring slots `batch` (host ints, repeats allowed) -> TransitionBatch in that order."""
        batch = np.asarray(batch, dtype=np.int64)
        gids = self.slot_graph_h[batch]
        if (gids < 0).any():
            raise IndexError('replay slot never written')
        ns = self.pool.ns[gids]
        dev = self.device
        # node offsets, adjacency offsets, vertex names and the slots themselves: one host->device copy
        node_off, adj_off, names, slots_d = TransitionBatch.index_to_device(ns, dev, extra=batch)
        n_vert, n_adj = int(ns.sum()), int((ns * ns).sum())
        adj, cur, nxt, action, reward, done = ops.replay_gather(
            slots_d.int(), self.ring, self.pool.adj, self.pool.adj_off, self.pool.n, self.pool.n_max, node_off,
            self._adj_off_plus(adj_off, n_adj), n_vert, n_adj, err=self._err)
        return TransitionBatch(ns, adj, cur, nxt, action, reward, done, index_dev=(node_off, adj_off, names))

    @staticmethod
    def _adj_off_plus(adj_off, n_adj):
        """gcrl_replay_gather takes [batch + 1] adjacency offsets; the kernel reads entries [0, batch) only (the last one is
        the total, which it gets as a size), so the [batch] view is passed with one padding element when it is empty."""
        return adj_off if adj_off.numel() else adj_off.new_zeros(1)

    def sample_buffer(self, sample_size):
        return self.gather(self.sample_indices(sample_size))

    def check(self):
        """Synchronising: raise if any gather since the last check saw offsets that disagree with the ring."""
        e = int(self._err.item())
        if e:
            self._err.zero_()
            raise RuntimeError('replay gather: host vertex counts disagree with the ring at minibatch position %d' % (e - 1))
