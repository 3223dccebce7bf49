"""Target of scripts/sanitize.sh (compute-sanitizer memcheck / racecheck / synccheck / initcheck): the hot path on
shapes that exercise every tile-boundary case -- tiny and ragged graphs (isolated vertices, 2-5 vertex graphs, dozens
of segment boundaries per tile) and one 200-vertex graph (199-edge segments spanning two tiles) -- forward + backward
in the math modes named on the command line, a short greedy rollout (environment kernels, score/argmax) and one Adam /
soft-update step.  Results are compared with the fp32 mode so that a tool-induced miscompute would show."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def er(n, p, rng):
    a = (rng.random((n, n)) < p).astype(np.uint8)
    a = np.triu(a, 1)
    return a | a.T


def main():
    modes = sys.argv[1].split(',') if len(sys.argv) > 1 else ['fp32', 'bf16x3', 'bf16']
    big = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    from graph_colouring_with_rl_b200.rollouts import greedy_rollouts
    from graph_colouring_with_rl_b200.trainer import QLearner, TransitionBatch, synthetic_transitions
    dev = torch.device('cuda:0')
    rng = np.random.default_rng(0)
    sizes = [2, 3, 5, 4, 2, 9, 17, 3, 33, 2, big]
    adjs = [er(n, 0.4, rng) for n in sizes]
    adjs[3][:] = 0                                   # a graph of isolated vertices
    ns, host = synthetic_transitions(adjs, 0.5, seed=1, device=dev)
    ref = None
    for mode in modes:
        torch.manual_seed(0)
        agent = DQNAgentGN(2, 1, 0, math_mode=mode)
        with torch.no_grad():
            agent.gn_behaviour.flat_params().mul_(1.5)
        agent.soft_update(agent.gn_behaviour, agent.gn_target, 1)
        learner = QLearner(agent, max_edges_per_micro_batch=2000, local_only=True)      # several micro-batches
        batch = TransitionBatch.from_host(ns, host, dev)
        loss = float(learner.learn_step(batch).item())
        g = agent.gn_behaviour._flat_grad.cpu()
        assert np.isfinite(loss) and bool(torch.isfinite(g).all())
        if ref is None:
            ref = (loss, g)
        else:
            tol = 5e-2 if mode == 'bf16x3' else 0.5
            assert abs(loss - ref[0]) <= tol * max(1.0, abs(ref[0])), (mode, loss, ref[0])
        out = greedy_rollouts(adjs[:6], agent.gn_behaviour, graph=False)
        used = out['colours_used'].cpu().numpy()
        assert used.min() >= 1
        print('sanitize target: mode %s loss %.6f |grad| %.4e colours %s' % (mode, loss, float(g.norm()), used.tolist()), flush=True)
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
