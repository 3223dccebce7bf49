"""Batched greedy colouring rollouts that never leave the device: first vertex by the
neighbour-count rule, then argmax-Q over uncoloured vertices until every graph is done
(test_policy_functions.py:36-73, batch of G graphs instead of one graph at a time).

Per step: gcrl_gn_forward (Q of every vertex of every graph; the vertex side of each block and the
fc1/fc2/fc3 head are one fused launch) -> gcrl_env_step_greedy (masked per-graph first-argmax fused
with the first-fit colouring, leaf pass and observation update).  Finished graphs get action -1
from the argmax (no uncoloured vertex) and are left alone.

Every step of an episode has the same shapes and the same buffers (only the colours change), so the
step -- about 30 kernel launches -- is captured ONCE into a CUDA graph and replayed: the host issues
one graph launch per step instead of walking the launch chain through Python and ctypes, which is
what bounded small shards (1 250 graphs per GPU at N = 8: 3.5 ms per step, a third of it launch gaps).
"""
import torch

from . import _lib, ops
from .envs import BatchedColouringEnv


class _GraphedStep(object):
    """One rollout step captured as a CUDA graph over an environment's fixed buffers."""

    def __init__(self, env, net):
        self.key = (net.flat_params().data_ptr(), net.math_mode, env.x.data_ptr())
        packed, eattr = env.packed
        flat = net.flat_params()
        dev = env.device
        lib = _lib.load()
        # a private workspace: the shared grow-only one may be re-allocated by a later, larger call
        self.ws = torch.empty(lib.gcrl_gn_forward_workspace_bytes(packed.n_vertices, packed.n_edges, 0), dtype=torch.uint8,
                              device=dev)
        self.graph = torch.cuda.CUDAGraph()
        # capture_begin / capture_end on a side stream instead of the `torch.cuda.graph` context manager: that one
        # synchronises, runs the garbage collector and EMPTIES the caching allocator on entry -- with tens of GB of
        # rollout workspaces cached, a per-environment capture then cost more than the sweep it speeds up
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self.graph.capture_begin()
            try:
                _, q, _ = ops.gn_forward(flat, net.dims, packed, env.x, eattr, math_mode=net.math_mode, ws=self.ws)
                self.action = env.step_greedy(q)
            finally:
                self.graph.capture_end()
        cur.wait_stream(side)

    def replay(self):
        self.graph.replay()
        return self.action


def _eager_step(env, net, packed, eattr, flat):
    _, q, _ = ops.gn_forward(flat, net.dims, packed, env.x, eattr, math_mode=net.math_mode)
    return env.step_greedy(q)


@torch.no_grad()
def greedy_rollouts(graphs, net, record=False, sync_every=8, env=None, graph=True):
    """graphs: list of adjacency matrices / target dicts (or pass a ready `env`).
    graph: replay the step as a CUDA graph (captured after one eager step, cached on the env).
    Returns dict(colours_used int32 [G] (device), steps, env[, actions list of int32 [G]])."""
    if env is None:
        env = BatchedColouringEnv(graphs, device=net.device)
    else:
        env.reset()
    packed, eattr = env.packed
    flat = net.flat_params()
    acts = []
    a = env.first_vertices()
    env.step(a)
    if record:
        acts.append(a)
    steps, max_steps = 1, int(env.ns.max())
    graphed = None
    if graph:
        graphed = getattr(env, '_graphed_step', None)
        if graphed is not None and graphed.key != (flat.data_ptr(), net.math_mode, env.x.data_ptr()):
            graphed = env._graphed_step = None
    while steps < max_steps:
        # the host looks at `done` every sync_every steps, and before each of the last few: leaf vertices colour themselves,
        # so the longest episode of a batch usually ends one or two steps short of n (one forward pass of the whole batch each)
        if (steps % sync_every == 0 or steps + 3 >= max_steps) and bool(env.done.all().item()):
            break
        if graphed is not None:
            a = graphed.replay()
            if record:
                a = a.clone()
        else:
            a = _eager_step(env, net, packed, eattr, flat)      # also the warm-up every capture needs
            if graph and steps + 2 < max_steps:
                graphed = env._graphed_step = _GraphedStep(env, net)
        if record:
            acts.append(a)
        steps += 1
    out = dict(colours_used=env.colours_used(), steps=steps, env=env)
    if record:
        out['actions'] = acts
    return out


def sequences_from_actions(actions):
    """list of [G] action tensors -> per-graph vertex sequences (dropping the -1 padding)."""
    a = torch.stack(actions, 0).cpu().numpy()
    return [[int(v) for v in a[:, g] if v >= 0] for g in range(a.shape[1])]
