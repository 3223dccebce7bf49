"""Large-minibatch DQN learning over complete-graph transitions (device-resident
replay data).  A `TransitionBatch` holds G transitions as flat arrays -- dense adjacency
bytes, colour vectors before/after the step, action, reward, done -- i.e. what
dqn_agent.py:147-223 consumes, without per-transition Python objects.  `QLearner` runs the
learn step over it in micro-batches sized to fit HBM (the edge state of 4096 graphs x 200
vertices is 41.7 GB per fp32 [E,64] tensor), accumulating one flat gradient, which at
world_size > 1 is summed -- together with the loss and the per-rank graph count, in the tail of the same buffer --
with a single asynchronous NCCL all-reduce (794 KB) that the fused Adam's stream waits on.
"""
import numpy as np
import torch

from . import _lib, ops

FIELDS = ('adj', 'cur_colour', 'next_colour', 'action', 'reward', 'done')


class TransitionBatch(object):
    def __init__(self, ns, adj, cur_colour, next_colour, action, reward, done, index_dev=None):
        """ns: host int64 [G] vertex counts; the rest are device tensors
        (adj uint8 [sum n^2], colours int32 [N], action int32 [G], reward f32 [G], done int32 [G]).
        index_dev: (node_off int32 [G+1], adj_off int64 [G], names f32 [N]) if the caller already has them on the device."""
        self.ns = np.asarray(ns, dtype=np.int64)
        self.adj, self.cur_colour, self.next_colour = adj, cur_colour, next_colour
        self.action, self.reward, self.done = action, reward, done
        dev = adj.device
        self.device = dev
        self.n_graphs = int(self.ns.size)
        self.node_off_h = np.concatenate([[0], np.cumsum(self.ns)]).astype(np.int64)
        self.adj_off_h = np.concatenate([[0], np.cumsum(self.ns * self.ns)]).astype(np.int64)
        self.edge_off_h = np.concatenate([[0], np.cumsum(self.ns * (self.ns - 1))]).astype(np.int64)
        self.n_vertices = int(self.node_off_h[-1])
        self.n_edges = int(self.edge_off_h[-1])
        if index_dev is not None:
            self.node_off, self.adj_off, self.names = index_dev
        elif self.n_vertices <= (1 << 16):
            # small batches: the index arrays travel in ONE host->device copy (a learn step at the reference's minibatch size
            # is bound by host-side issue time: every separate pageable copy is a blocking call)
            self.node_off, self.adj_off, self.names = self.index_to_device(self.ns, dev)[:3]
        else:
            self.node_off = torch.from_numpy(self.node_off_h.astype(np.int32)).to(dev)
            self.adj_off = torch.from_numpy(self.adj_off_h[:-1].copy()).to(dev)
            counts = torch.from_numpy(self.ns).to(dev)
            start = torch.repeat_interleave(self.node_off[:-1].long(), counts, output_size=self.n_vertices)   # (no sync)
            self.names = (torch.arange(self.n_vertices, device=dev) - start).float()   # x[:,0] (graph_colouring_env.py:124-128)

    @staticmethod
    def index_to_device(ns, dev, extra=None):
        """[node_off | adj_off | vertex names | extra] as ONE int64 host->device copy -> (node_off int32 [G+1], adj_off int64
        [G], names f32 [N], extra int64 or None)."""
        ns = np.asarray(ns, dtype=np.int64)
        G, N = int(ns.size), int(ns.sum())
        node_off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
        adj_off = np.concatenate([[0], np.cumsum(ns * ns)[:-1]]).astype(np.int64) if G else np.zeros(0, dtype=np.int64)
        names = np.arange(N, dtype=np.int64) - np.repeat(node_off[:-1], ns)
        parts = [node_off, adj_off, names] + ([np.asarray(extra, dtype=np.int64)] if extra is not None else [])
        d = torch.from_numpy(np.concatenate(parts)).to(dev, non_blocking=True)
        return (d[:G + 1].int(), d[G + 1:2 * G + 1], d[2 * G + 1:2 * G + 1 + N].float(),
                d[2 * G + 1 + N:] if extra is not None else None)

    @classmethod
    def empty(cls, device):
        """A minibatch without graphs (a rank whose replay ring is still empty joins the collective with this)."""
        dev = torch.device(device)
        z = lambda dt: torch.zeros(0, dtype=dt, device=dev)
        return cls(np.zeros(0, dtype=np.int64), z(torch.uint8), z(torch.int32), z(torch.int32), z(torch.int32), z(torch.float32),
                   z(torch.int32))

    @classmethod
    def from_host(cls, ns, host, device):
        """host: dict of (ideally pinned) CPU tensors named like FIELDS; async H2D copies."""
        dev = torch.device(device)
        t = {k: host[k].to(dev, non_blocking=True) for k in FIELDS}
        return cls(ns, t['adj'], t['cur_colour'], t['next_colour'], t['action'], t['reward'], t['done'])

    @staticmethod
    def host_bytes(host):
        return int(sum(host[k].numel() * host[k].element_size() for k in FIELDS))

    def x_of(self, colour, v0=0, v1=None):
        """Observation x = [vertex name, colour] of vertices [v0, v1) given their colours."""
        v1 = self.n_vertices if v1 is None else v1
        return torch.stack([self.names[v0:v1], colour.float()], 1)

    def micro_ranges(self, max_edges):
        """Contiguous graph ranges whose complete edge count stays within max_edges."""
        out, g0 = [], 0
        while g0 < self.n_graphs:
            g1 = g0 + 1
            while g1 < self.n_graphs and self.edge_off_h[g1 + 1] - self.edge_off_h[g0] <= max_edges:
                g1 += 1
            out.append((g0, g1))
            g0 = g1
        return out


class _Slice(object):
    """Views of one micro-batch [g0, g1) of a TransitionBatch."""

    def __init__(self, b, g0, g1):
        v0, v1 = int(b.node_off_h[g0]), int(b.node_off_h[g1])
        a0, a1 = int(b.adj_off_h[g0]), int(b.adj_off_h[g1])
        self.v0, self.v1, self.g0, self.g1 = v0, v1, g0, g1
        self.n_vertices = v1 - v0
        self.n_edges = int(b.edge_off_h[g1] - b.edge_off_h[g0])
        self.adj = b.adj[a0:a1]
        self.adj_off = b.adj_off[g0:g1] - a0
        self.node_off = b.node_off[g0:g1 + 1] - v0
        self.packed, self.eattr = ops.pack_dense(self.adj, self.adj_off, self.node_off, self.n_vertices, self.n_edges)


def reduce_mean_gradient(gbuf, n_params, group=None, batch_total=None):
    """The one collective of a multi-GPU learn step.  gbuf = [flat gradient (n_params) | loss | graph count]
    holds this rank's SUMS.  With `batch_total` given, the sums were already scaled by 1/batch_total (the caller knows
    the global minibatch size) and the all-reduce finishes the job.  Without it every rank contributed unscaled sums
    and its own graph count -- ranks may hold different numbers of graphs (shard_range(7, r, 4) gives 1, 2, 2, 2) -- and
    gradient and loss are divided by the all-reduced count on the device, so the result is the single-process mean
    over the union minibatch (dqn_agent.py:215: mse_loss over the whole batch) whatever the split.  The collective is
    issued asynchronously; the consumer's stream waits for it, the host does not."""
    work = torch.distributed.all_reduce(gbuf, group=group, async_op=True)
    work.wait()
    if batch_total is None:
        gbuf[:n_params + 1].div_(gbuf[n_params + 1])
    return gbuf


class QLearner(object):
    def __init__(self, agent, max_edges_per_micro_batch=512 * 200 * 199, process_group=None, local_only=False,
                 sync_params=True):
        """local_only: never communicate, even inside an initialised process group (a learner that only some ranks
        run, e.g. the single-GPU training loop, must not enter collectives the other ranks do not).
        sync_params: at world_size > 1 broadcast rank 0's behaviour and target parameters (and Adam state) once, so
        that the replicas are identical whatever each rank's seed was."""
        self.agent = agent
        self.max_edges = int(max_edges_per_micro_batch)
        self.group = process_group
        self.world = 1
        if not local_only and (process_group is not None or
                               (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.world = torch.distributed.get_world_size(process_group)
        self._stash = None
        self._side = None                # side stream + private forward workspace of the target forward (small minibatches)
        self._side_ws = None
        self._gbuf = None
        self.keep_q = False              # tests: keep (q_expected, q_target) of the last fwd_bwd
        self.last_q = None
        if self.world > 1 and sync_params and hasattr(agent, 'gn_behaviour'):
            self.broadcast_params()

    def broadcast_params(self, src=0):
        agent = self.agent
        nets = [agent.gn_behaviour] + ([agent.gn_target] if hasattr(agent, 'gn_target') else [])
        for net in nets:
            torch.distributed.broadcast(net.flat_params(), src=src, group=self.group)
        opt = agent.gn_behaviour.optimizer
        if opt.exp_avg is not None:
            torch.distributed.broadcast(opt.exp_avg, src=src, group=self.group)
            torch.distributed.broadcast(opt.exp_avg_sq, src=src, group=self.group)

    def _grad_buffer(self):
        """[P + 2] floats: the behaviour net's flat gradient, then the loss and this rank's graph count -- one buffer,
        one all-reduce.  The net's `_flat_grad` (what FusedAdam reads) is the leading view."""
        net = self.agent.gn_behaviour
        flat = net.flat_params(full_check=False)
        P = flat.numel()
        g = self._gbuf
        if g is None or g.device != flat.device or g.numel() != P + 2 or net._flat_grad is None \
                or net._flat_grad.data_ptr() != g.data_ptr():
            g = self._gbuf = torch.zeros(P + 2, dtype=torch.float32, device=flat.device)
            net._flat_grad = g[:P]
        return g, P

    @torch.no_grad()
    def target_q(self, batch, slices=None, ws=None):
        """Target-network Q of every vertex of the next states (dqn_agent.py:187-188) -> [N].
        slices: micro-batch layouts already built for this batch (learn_step shares them between its two forwards)."""
        net = self.agent.gn_target
        flat = net.flat_params(full_check=False)
        out = torch.empty(batch.n_vertices, dtype=torch.float32, device=batch.device)
        for i, (g0, g1) in enumerate(batch.micro_ranges(self.max_edges)):
            s = slices[i] if slices is not None else _Slice(batch, g0, g1)
            x = batch.x_of(batch.next_colour[s.v0:s.v1], s.v0, s.v1)
            _, q, _ = ops.gn_forward(flat, net.dims, s.packed, x, s.eattr, math_mode=net.math_mode, ws=ws)
            out[s.v0:s.v1] = q
        return out

    @torch.no_grad()
    def fwd_bwd(self, batch, q_next, batch_total=None, slices=None, q_next_ready=None):
        """Behaviour forward, TD loss against q_next, backward.  Leaves the mean gradient over the GLOBAL minibatch
        (all ranks' graphs) in the behaviour net's flat gradient buffer and returns the global loss [1] (a view of the
        reduction buffer: read or clone it before the next call).  batch_total: the global number of graphs if the
        caller knows it; otherwise it is all-reduced together with the gradient (see reduce_mean_gradient)."""
        agent = self.agent
        net = agent.gn_behaviour
        flat = net.flat_params(full_check=False)
        gbuf, P = self._grad_buffer()
        grad, loss = gbuf[:P], gbuf[P:P + 1]
        gbuf[P:].zero_()
        if self.world == 1:
            total = batch.n_graphs if batch_total is None else int(batch_total)
        else:
            total = 1 if batch_total is None else int(batch_total)     # unscaled sums, normalised after the reduce
            gbuf[P + 1] = float(batch.n_graphs)
        first = True
        kept = []
        for i, (g0, g1) in enumerate(batch.micro_ranges(self.max_edges)):
            s = slices[i] if slices is not None else _Slice(batch, g0, g1)
            x = batch.x_of(batch.cur_colour[s.v0:s.v1], s.v0, s.v1)
            _, q, self._stash = ops.gn_forward(flat, net.dims, s.packed, x, s.eattr, with_stash=True,
                                               math_mode=net.math_mode, stash=self._stash)
            if q_next_ready is not None:              # q_next was computed on a side stream (learn_step)
                torch.cuda.current_stream(batch.device).wait_event(q_next_ready)
                q_next_ready = None
            _, q_exp, q_tgt, d_q = ops.td_target_loss(q, q_next[s.v0:s.v1], batch.next_colour[s.v0:s.v1], s.node_off,
                                                      batch.action[g0:g1], batch.reward[g0:g1], batch.done[g0:g1],
                                                      agent.gamma, batch_total=total, loss_out=loss)
            if self.keep_q:
                kept.append((q_exp, q_tgt))
            ops.gn_backward(flat, net.dims, s.packed, x, s.eattr, self._stash, d_q=d_q, grad=grad,
                            accumulate=not first, math_mode=net.math_mode)
            first = False
        if first:
            grad.zero_()                                                # a rank without graphs still joins the reduce
        if self.keep_q:
            self.last_q = (torch.cat([k[0] for k in kept]), torch.cat([k[1] for k in kept])) if kept else None
        if self.world > 1:
            reduce_mean_gradient(gbuf, P, self.group, batch_total)
        return loss

    @torch.no_grad()
    def learn_step(self, batch, batch_total=None):
        """dqn_agent.py:147-223 on a TransitionBatch: target forward, behaviour forward/backward,
        fused Adam + soft target update.  The micro-batch layouts are built once for both forwards.  A minibatch that is a
        single micro-batch (the reference's 64 transitions: launch chains of ~10 us kernels that leave most SMs idle) runs
        its target forward on a side stream next to the behaviour forward; the TD loss waits for both."""
        agent = self.agent
        agent.gn_behaviour.flat_params()             # one full aliasing check per step; the inner calls use the cheap one
        agent.gn_target.flat_params()
        ranges = batch.micro_ranges(self.max_edges)
        slices = [_Slice(batch, g0, g1) for g0, g1 in ranges] if len(ranges) <= 2 else None
        if len(ranges) == 1 and batch.n_edges <= (1 << 21):
            dev = batch.device
            main = torch.cuda.current_stream(dev)
            if self._side is None:
                self._side = torch.cuda.Stream(dev)
            nb = _lib.load().gcrl_gn_forward_workspace_bytes(batch.n_vertices, batch.n_edges, 0)
            if self._side_ws is None or self._side_ws.numel() < nb:
                self._side_ws = torch.empty(max(int(nb), 256), dtype=torch.uint8, device=dev)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                q_next = self.target_q(batch, slices, ws=self._side_ws)
                ready = torch.cuda.Event()
                ready.record(self._side)
            q_next.record_stream(main)
            loss = self.fwd_bwd(batch, q_next, batch_total, slices, q_next_ready=ready)
        else:
            q_next = self.target_q(batch, slices)
            loss = self.fwd_bwd(batch, q_next, batch_total, slices)
        agent.gn_behaviour.optimizer.step(target_flat=agent.gn_target.flat_params(full_check=False), tau=agent.tau,
                                          grad=agent.gn_behaviour._flat_grad)
        agent.last_loss = loss.clone()
        return agent.last_loss


@torch.no_grad()
def synthetic_transitions(adjs, frac_coloured, seed, device):
    """Mid-episode transitions on the given graphs: colour about frac_coloured of each
    graph's vertices in uniformly random order through the device environment, then take one
    more random action and record (s, a, r, s', done).  Returns (ns, dict of CPU tensors)."""
    from .envs import BatchedColouringEnv
    env = BatchedColouringEnv(adjs, device=device)
    gen = torch.Generator(device=env.device)
    gen.manual_seed(seed)

    def random_actions(active):
        r = torch.rand(env.n_vertices, generator=gen, device=env.device)
        a = ops.score_argmax(r, env.colour, env.node_off)
        return torch.where(active, a, torch.full_like(a, -1))
    steps = int(frac_coloured * int(env.ns.max()))
    target = torch.from_numpy((frac_coloured * env.ns).astype(np.int64)).to(env.device)
    counts = torch.from_numpy(env.ns).to(env.device)
    gid = torch.repeat_interleave(torch.arange(env.n_graphs, device=env.device), counts, output_size=env.n_vertices)
    for _ in range(steps):
        coloured = torch.zeros(env.n_graphs, dtype=torch.long, device=env.device).index_add_(0, gid, (env.colour >= 0).long())
        uncol = counts - coloured
        env.step(random_actions((coloured < target) & (uncol > 1)))
    cur = env.colour.clone()
    act = random_actions(torch.ones(env.n_graphs, dtype=torch.bool, device=env.device))
    env.step(act)
    host = dict(adj=env.adj.cpu(), cur_colour=cur.cpu(), next_colour=env.colour.cpu(), action=act.cpu(),
                reward=env.reward.cpu(), done=env.done.cpu())
    return env.ns.copy(), host


def shard_host(ns, host, lo, hi):
    """Graphs [lo, hi) of a host transition dict (FIELDS) -> (ns[lo:hi], dict of views; pinned stays pinned)."""
    ns = np.asarray(ns, dtype=np.int64)
    node_off = np.concatenate([[0], np.cumsum(ns)])
    adj_off = np.concatenate([[0], np.cumsum(ns * ns)])
    v0, v1, a0, a1 = int(node_off[lo]), int(node_off[hi]), int(adj_off[lo]), int(adj_off[hi])
    return ns[lo:hi], dict(adj=host['adj'][a0:a1], cur_colour=host['cur_colour'][v0:v1],
                           next_colour=host['next_colour'][v0:v1], action=host['action'][lo:hi],
                           reward=host['reward'][lo:hi], done=host['done'][lo:hi])


def shard_batch(ns, host, lo, hi, device):
    """This rank's contiguous slice of a global minibatch as a device TransitionBatch (split by graph)."""
    return TransitionBatch.from_host(*shard_host(ns, host, lo, hi), device)


def shard_range(n_items, rank, world):
    """Contiguous split of n_items units (graphs) over `world` ranks -> [lo, hi) of `rank`."""
    return n_items * rank // world, n_items * (rank + 1) // world
