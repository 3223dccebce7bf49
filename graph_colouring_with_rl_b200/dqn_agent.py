"""Drop-in for the reference's dqn_agent.py: `ReplayBuffer` and `DQNAgentGN` with the same
constructor arguments, hyper-parameter fields and method signatures; act / learn run on the
device through libgcrl_b200.

Reference: dqn_agent.py:8-65 (ReplayBuffer), :68-238 (DQNAgentGN).

Differences that are deliberate and invisible to callers:
  * `learn()` makes no per-graph Python loops and no host round trips: two batched forwards,
    `gcrl_td_target_loss`, `gcrl_gn_backward`, and one fused Adam + soft-update launch;
  * the replay buffer keeps each transition's tensors where they already are (device tensors
    from the device environment stay on the device; the reference forces `.cpu()`), and when a
    transition carries a dense adjacency (`ColouringData`) the minibatch is batched straight
    from adjacency bytes with `gcrl_pack_dense` instead of concatenating edge lists;
  * random numbers are drawn from the same generators in the same order (`random.random`,
    `random.choice`, `np.random.choice`) so seeded runs take the same branches.
"""
import random
from collections import OrderedDict

import numpy as np
import torch

from . import _lib, ops
from .architecture import GNNetwork
from .data import Batch


class ReplayBuffer(object):
    def __init__(self, max_size, global_features_size):
        self.max_mem_size = max_size
        self.current_mem_size = 0
        self.data_memory = [None] * self.max_mem_size
        self.current_global_features = np.zeros((self.max_mem_size, global_features_size), dtype=np.float32)
        self.assigned_vertices_memory = np.zeros(self.max_mem_size, dtype=object)
        self.action_ind_memory = np.zeros((self.max_mem_size, 1))
        self.reward_memory = np.zeros(self.max_mem_size)
        self.next_data_memory = [None] * self.max_mem_size
        self.next_global_features = np.zeros((self.max_mem_size, global_features_size), dtype=np.float32)
        self.done_memory = np.zeros(self.max_mem_size, dtype=np.intc)

    def store_transition(self, data, current_global_features, assigned_vertices, action_ind, reward, next_data,
                         next_global_features, done):
        index = self.current_mem_size % self.max_mem_size          # dqn_agent.py:33
        self.data_memory[index] = data
        self.current_global_features[index] = current_global_features
        self.assigned_vertices_memory[index] = assigned_vertices
        self.action_ind_memory[index] = action_ind
        self.reward_memory[index] = reward
        self.next_data_memory[index] = next_data
        self.next_global_features[index] = next_global_features
        self.done_memory[index] = done
        self.current_mem_size += 1

    def sample_indices(self, sample_size):
        max_mem = min(self.current_mem_size, self.max_mem_size)
        return np.random.choice(max_mem, sample_size)              # with replacement (:48)

    def sample_buffer(self, sample_size):
        """Same 8-tuple as the reference (:45-65)."""
        batch = self.sample_indices(sample_size)
        return self.gather(batch)

    def gather(self, batch):
        datas_batch = Batch.from_data_list([self.data_memory[i] for i in batch])
        next_datas_batch = Batch.from_data_list([self.next_data_memory[i] for i in batch])
        return (datas_batch, self.current_global_features[batch], self.assigned_vertices_memory[batch],
                self.action_ind_memory[batch], self.reward_memory[batch], next_datas_batch,
                self.next_global_features[batch], self.done_memory[batch])


class DQNAgentGN():
    def __init__(self, len_vertex_features, len_edge_features, len_global_features, test_mode=False,
                 math_mode='fp32'):
        self.gamma = 1                 # dqn_agent.py:80-88
        self.alpha = 1e-3
        self.tau = 1e-3
        self.no_episodes = 25000
        self.eps_start = .9
        self.eps_end = .01
        self.max_buffer_size = int(1e5)
        self.batch_size = 64
        self.update_every = 5
        self.len_global_features = len_global_features
        self.learn_counter = 0
        self.gn_behaviour = GNNetwork(len_vertex_features, len_edge_features, len_global_features, name='GN',
                                      alpha=self.alpha, math_mode=math_mode)
        if not test_mode:
            self.gn_target = GNNetwork(len_vertex_features, len_edge_features, len_global_features,
                                       name='GN_target', alpha=self.alpha, math_mode=math_mode)
            self.soft_update(self.gn_behaviour, self.gn_target, 1)
            self.memory = ReplayBuffer(self.max_buffer_size, len_global_features)
        self.episode_no = 0
        self.last_loss = None          # device scalar of the most recent learn() (not in the reference)
        self._node_off1, self._node_off1_n = None, -1
        self.act_graphs = True         # batch-1 choose_action replays a CUDA graph per (graph, network); False: eager launches
        self._act_cache = OrderedDict()

    # ---- act (dqn_agent.py:108-140) ---------------------------------------------------------------
    def choose_action(self, data, global_features, unassigned_vertices, epsilon=None, test_mode=False):
        if len(unassigned_vertices) == 0:
            # the reference raises here too (np.argmax of an empty array / random.choice of an empty list)
            raise ValueError('choose_action: no unassigned vertex to choose from')
        if epsilon is None:
            if test_mode:
                epsilon = 0
            else:
                epsilon_decay = (self.eps_end / self.eps_start) ** (1 / self.no_episodes)
                epsilon = self.eps_start * (epsilon_decay ** self.episode_no)
        if random.random() > epsilon:
            net = self.gn_behaviour
            gn_input = data.to(net.device)
            n = gn_input.x.shape[0]
            if (self.act_graphs and net.dims == (2, 1, 0) and getattr(gn_input, '_unassigned', None) is unassigned_vertices
                    and gn_input.x.is_cuda):
                # an observation of this package's environment, scored for the list it was built from: the whole act path
                # (forward pass + masked first-argmax) replays as one CUDA graph per (graph, network)
                return self._graphed_act(gn_input, n)
            gf = torch.tensor(global_features, dtype=torch.float).unsqueeze(0).to(net.device)
            with torch.no_grad():
                _, action_values = net(gn_input, gf, [0] * n, None)
            vertex_ind = self._masked_argmax(action_values.view(-1), gn_input, unassigned_vertices, n)
        else:
            vertex_ind = random.choice(unassigned_vertices)
        return vertex_ind

    @torch.no_grad()
    def _graphed_act(self, data, n):
        """Batch-1 greedy action (dqn_agent.py:119-134) as ONE CUDA-graph replay: the environment hands out the same static
        edge tensors for every step of every episode on a graph, so the launch chain of the forward pass (block 1 .. block 5,
        head) and the masked first-argmax is captured once per (graph, network buffers, math mode) over a private workspace
        and a static copy of the observation; a call is then one 8 n-byte copy, one graph launch and the `.item()` that returns
        the Python int (the launch chain walked through Python and ctypes is ~0.25 ms for a 50-vertex graph).  Parameters
        are read in place from the network's flat buffer, so learning between calls needs no re-capture; a network whose
        buffer moved (`.to()`, `load_state_dict(assign=True)`) gets a new capture.  At most 64 graphs are kept."""
        from .architecture import _pack_cache
        net = self.gn_behaviour
        flat = net.flat_params()
        key = (data.edge_index.data_ptr(), n, flat.data_ptr(), net.math_mode)
        cache = self._act_cache
        ent = cache.get(key)
        if ent is not None and ent['edge_index'] is data.edge_index:
            cache.move_to_end(key)
            ent['x'].copy_(data.x)
            ent['graph'].replay()
            return int(ent['action'].item())
        dev = flat.device
        x_s = data.x.float().clone()
        packed, eattr_i = _pack_cache.get(data.edge_index, n, data.edge_attr.float())
        node_off = torch.tensor([0, n], dtype=torch.int32, device=dev)
        lib = _lib.load()
        ws = torch.empty(lib.gcrl_gn_forward_workspace_bytes(packed.n_vertices, packed.n_edges, 0), dtype=torch.uint8, device=dev)

        def body():
            _, q, _ = ops.gn_forward(flat, net.dims, packed, x_s, eattr_i, math_mode=net.math_mode, ws=ws)
            return ops.score_argmax(q, x_s[:, 1].to(torch.int32), node_off)   # -1 = uncoloured (graph_colouring_env.py:124-128)
        action = body()                                   # eager: the answer of this call, and the warm-up every capture needs
        out = int(action.item())
        graph = torch.cuda.CUDAGraph()
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            graph.capture_begin()
            try:
                action = body()
            finally:
                graph.capture_end()
        cur.wait_stream(side)
        cache[key] = dict(graph=graph, x=x_s, action=action, ws=ws, packed=packed, eattr=eattr_i, node_off=node_off,
                          edge_index=data.edge_index)
        while len(cache) > 64:
            cache.popitem(last=False)
        return out

    def _masked_argmax(self, q, data, unassigned_vertices, n):
        """np.argmax over q[unassigned_vertices] (dqn_agent.py:131-134) as the device's masked first-argmax.
        np.argmax returns the first maximum IN LIST ORDER; the device kernel returns the lowest vertex index among
        the maxima, which is the same thing when the list is ascending (it always is in the reference's callers:
        graph_colouring_env.py:152 rebuilds it in vertex order).  A list that is not ascending takes the literal
        reference arithmetic on the host.  Fast path: an observation from this package's environment carries the
        very list object it was built from (`GenerateData_v1` records it), so when the caller passes that list --
        the reference loop does, dqn_main.py:176 -- the mask is x[:,1] < 0 on the device: no numpy mask, no extra
        H2D copy; the one unavoidable sync is the `.item()` that returns the Python int."""
        dev = q.device
        if self._node_off1 is None or self._node_off1_n != n or self._node_off1.device != dev:
            self._node_off1 = torch.tensor([0, n], dtype=torch.int32, device=dev)
            self._node_off1_n = n
        if getattr(data, '_unassigned', None) is unassigned_vertices and data.x.shape[1] >= 2:
            colour = data.x[:, 1].to(torch.int32)              # -1 = uncoloured (graph_colouring_env.py:124-128)
            return int(ops.score_argmax(q, colour, self._node_off1).item())
        un = np.asarray(unassigned_vertices, dtype=np.int64)
        if un.size > 1 and np.any(un[1:] <= un[:-1]):
            return int(un[int(np.argmax(q[torch.from_numpy(un).to(dev)].cpu().numpy()))])
        mask = np.zeros(n, dtype=np.int32)
        mask[un] = -1
        return int(ops.score_argmax(q, torch.from_numpy(mask).to(dev), self._node_off1).item())

    def remember(self, data, current_global_features, assigned_vertices, action_ind, reward, next_data,
                 next_global_features, done):
        self.memory.store_transition(data, current_global_features, assigned_vertices, action_ind, reward,
                                     next_data, next_global_features, done)

    # ---- learn (dqn_agent.py:147-223) ---------------------------------------------------------------
    def _dense_ok(self, datas):
        """The adjacency fast path batches from dense bytes and builds edge_attr = -adj itself, so it is only valid
        for observations of this package's environment (ColouringData: complete graph, one static -1/0 edge
        feature).  Anything else -- a caller's own Data with other edge attributes or len_edge_features != 1 --
        goes through the generic edge-list path."""
        from .envs import ColouringData
        return self.gn_behaviour.dims[1] == 1 and all(
            isinstance(d, ColouringData) and getattr(d, 'adj', None) is not None for d in datas)

    def _collate(self, datas, device):
        """-> (packed, eattr internal, x [N,vd], node_off int32 [G+1])."""
        ns = [int(d.x.shape[0]) for d in datas]
        node_off = torch.tensor(np.concatenate([[0], np.cumsum(ns)]), dtype=torch.int32, device=device)
        x = torch.cat([d.x.to(device, torch.float32) for d in datas], 0)
        if self._dense_ok(datas):
            adj = torch.cat([d.adj.to(device).reshape(-1) for d in datas])
            nn_ = np.asarray(ns, dtype=np.int64)
            adj_off = torch.tensor(np.concatenate([[0], np.cumsum(nn_ * nn_)[:-1]]), dtype=torch.int64, device=device)
            packed, eattr = ops.pack_dense(adj, adj_off, node_off, int(nn_.sum()), int((nn_ * (nn_ - 1)).sum()))
            return packed, eattr, x, node_off
        b = Batch.from_data_list([d if d.x.device == d.edge_index.device else d.to(device) for d in datas])
        ei = b.edge_index.to(device)
        packed = ops.pack(ei, x.shape[0])
        eattr = ops.permute_rows(b.edge_attr.to(device, torch.float32).reshape(ei.shape[1], -1), packed.perm, True)
        return packed, eattr, x, node_off

    @staticmethod
    def _same_edges(cur, nxt):
        """True when every next-state observation shares its edge tensors with the current one (the environment's
        static per-graph tensors: graph_colouring_env.py:136-137 hands out target['edge_list'] / ['edge_attr'] every
        step), so one batched layout serves both forwards."""
        for c, n in zip(cur, nxt):
            if c.x.shape[0] != n.x.shape[0]:
                return False
            for k in ('edge_index', 'edge_attr'):
                a, b = getattr(c, k), getattr(n, k)
                if a is b:
                    continue
                if a is None or b is None or a.shape != b.shape or a.data_ptr() != b.data_ptr():
                    return False
        return True

    def learn(self):
        if self.memory.current_mem_size < 10 * self.batch_size:
            return
        if self.learn_counter == 0:
            print('Starting learning')
            self.learn_counter += 1
        mem = self.memory
        batch = mem.sample_indices(self.batch_size)
        self.learn_on(batch)

    def learn_on(self, batch):
        """One DQN step on the replay slots `batch` (what learn() does after sampling)."""
        mem = self.memory
        net_b, net_t = self.gn_behaviour, self.gn_target
        dev = net_b.device
        cur = [mem.data_memory[i] for i in batch]
        nxt = [mem.next_data_memory[i] for i in batch]
        packed, eattr, x_cur, node_off = self._collate(cur, dev)
        if self._same_edges(cur, nxt):
            packed_n, eattr_n = packed, eattr
            x_next = torch.cat([d.x.to(dev, torch.float32) for d in nxt], 0)
        else:       # the reference batches next_datas separately (dqn_agent.py:55-56,187)
            packed_n, eattr_n, x_next, _ = self._collate(nxt, dev)
        # assigned-after-step one-hot (dqn_main.py:184) -> "colour >= 0 means assigned"
        assigned = np.concatenate([np.asarray(mem.assigned_vertices_memory[i]).reshape(-1) for i in batch])
        next_colour = torch.from_numpy(np.where(assigned == 1, 0, -1).astype(np.int32)).to(dev)
        action_h = mem.action_ind_memory[batch].reshape(-1).astype(np.int64)
        n_h = np.asarray([int(d.x.shape[0]) for d in cur], dtype=np.int64)
        if np.any(action_h < 0) or np.any(action_h >= n_h):
            raise IndexError('replay transition with an action outside its graph: %r' % (action_h[(action_h < 0) | (action_h >= n_h)][:4],))
        action = torch.from_numpy(action_h.astype(np.int32)).to(dev)
        reward = torch.from_numpy(mem.reward_memory[batch].astype(np.float32)).to(dev)
        done = torch.from_numpy(mem.done_memory[batch].astype(np.int32)).to(dev)
        gfeat = gfeat_n = vgraph = None
        if self.len_global_features > 0:
            gfeat = torch.from_numpy(mem.current_global_features[batch]).to(dev)
            gfeat_n = torch.from_numpy(mem.next_global_features[batch]).to(dev)
            counts = (node_off[1:] - node_off[:-1]).long()
            vgraph = torch.repeat_interleave(torch.arange(len(batch), device=dev), counts).int()
        flat_b, flat_t = net_b.flat_params(), net_t.flat_params()
        _, q_cur, stash = ops.gn_forward(flat_b, net_b.dims, packed, x_cur, eattr, gfeat, vgraph, with_stash=True,
                                         math_mode=net_b.math_mode)
        _, q_next, _ = ops.gn_forward(flat_t, net_t.dims, packed_n, x_next, eattr_n, gfeat_n, vgraph,
                                      math_mode=net_t.math_mode)
        loss, q_exp, q_tgt, d_q = ops.td_target_loss(q_cur, q_next, next_colour, node_off, action, reward, done,
                                                     self.gamma)
        if net_b._flat_grad is None:
            net_b._flat_grad = torch.zeros_like(flat_b)
        grad = ops.gn_backward(flat_b, net_b.dims, packed, x_cur, eattr, stash, d_q=d_q, gfeat=gfeat, vgraph=vgraph,
                               grad=net_b._flat_grad, accumulate=False, math_mode=net_b.math_mode)
        # optimizer.step() and soft_update(behaviour, target, tau) fused (dqn_agent.py:220,223)
        net_b.optimizer.step(target_flat=flat_t, tau=self.tau, grad=grad)
        self.last_loss = loss
        return loss

    def soft_update(self, behaviour_network, target_network, tau):
        ops.soft_update(behaviour_network.flat_params(), target_network.flat_params(), tau)

    def save_models(self, checkpoint_filepath_root):
        return self.gn_behaviour.save_checkpoint(checkpoint_filepath_root)

    def load_models(self, checkpoint_filepath_root):
        self.gn_behaviour.load_checkpoint(checkpoint_filepath_root)
