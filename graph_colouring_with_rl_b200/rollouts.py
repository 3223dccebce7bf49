"""Batched greedy colouring rollouts that never leave the device: first vertex by the
neighbour-count rule, then argmax-Q over uncoloured vertices until every graph is done
(test_policy_functions.py:36-73, batch of G graphs instead of one graph at a time).

Per step: gcrl_gn_forward (Q of every vertex of every graph) -> gcrl_score_argmax (masked
per-graph first-argmax) -> gcrl_env_step (first-fit colour, leaf pass, observation update).
Finished graphs get action -1 from the argmax (no uncoloured vertex) and are left alone.
"""
import torch

from . import ops
from .envs import BatchedColouringEnv


@torch.no_grad()
def greedy_rollouts(graphs, net, record=False, sync_every=8, env=None):
    """graphs: list of adjacency matrices / target dicts (or pass a ready `env`).
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
    while steps < max_steps:
        if steps % sync_every == 0 and bool(env.done.all().item()):
            break
        _, q, _ = ops.gn_forward(flat, net.dims, packed, env.x, eattr, math_mode=net.math_mode)
        a = ops.score_argmax(q, env.colour, env.node_off)
        env.step(a)
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
