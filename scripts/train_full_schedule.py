"""End-to-end check that the whole pipeline LEARNS: the reference's full training schedule (dqn_agent.py:80-88: 25 000
episodes, eps 0.9 -> 0.01, replay 1e5, minibatch 64, learn every 5 environment steps, Adam 1e-3, tau 1e-3) run by the device
training loop (dqn_main.DeviceTrainer, 64 episodes in lock step) on 1000 synthetic training graphs of 15-50 vertices
(the reference's training_dataset.pickle is not in the checkout: Erdos-Renyi graphs with p ~ U(0.1, 0.7) and
Watts-Strogatz / Barabasi-Albert graphs from networkx stand in for its seven families), validated every 2 500 episodes on
the reference's OWN 100 validation graphs (tests/golden/validation_graphs.npz), whose DSATUR colour counts (mean 7.42) and
random-order counts (mean 8.58) are the yardsticks.

    python scripts/train_full_schedule.py [episodes] [math_mode] [weights_root] > profiles/r02_train_full_schedule.json
"""
import contextlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]


def training_graphs(rng, count=1000):
    import networkx as nx
    out = []
    for i in range(count):
        n = int(rng.integers(15, 51))
        kind = i % 4
        if kind < 2:
            p = rng.uniform(0.1, 0.7)
            a = np.triu((rng.random((n, n)) < p).astype(np.uint8), 1)
            a = a | a.T
        elif kind == 2:
            g = nx.watts_strogatz_graph(n, int(rng.integers(2, max(3, n // 3))) // 2 * 2 + 2, rng.uniform(0.1, 0.5), seed=int(rng.integers(1 << 30)))
            a = nx.to_numpy_array(g, dtype=np.uint8)
        else:
            g = nx.barabasi_albert_graph(n, int(rng.integers(1, max(2, n // 4))), seed=int(rng.integers(1 << 30)))
            a = nx.to_numpy_array(g, dtype=np.uint8)
        out.append(np.ascontiguousarray(a))
    return out


def main():
    episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
    math_mode = sys.argv[2] if len(sys.argv) > 2 else 'bf16x3'
    import gcrl_testutil as util
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    from graph_colouring_with_rl_b200.dqn_main import DeviceTrainer
    val = util.load_validation_graphs()
    val_targets = [dict(adj=g['adj'], name=g['name']) for g in val]
    dsatur = float(np.mean([g['dsatur_colours'] for g in val]))
    rng = np.random.default_rng(0)
    np.random.seed(0)
    torch.manual_seed(0)
    train = training_graphs(rng)
    agent = DQNAgentGN(2, 1, 0, math_mode=math_mode)
    agent.no_episodes = episodes                          # the epsilon schedule spans the run (dqn_agent.py:112-117)
    with contextlib.redirect_stdout(sys.stderr):
        tr = DeviceTrainer(agent, train, val_targets, n_envs=64, device_seed=0, validate_every=2500)
        t0 = time.perf_counter()
        tr.run(episodes, verbose=True, log_every=2500)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if len(sys.argv) > 3:                                 # save the trained behaviour network (the reference's <root>_GN state_dict blob)
        with contextlib.redirect_stdout(sys.stderr):
            agent.save_models(sys.argv[3])
    curve = {int(ep): float(np.mean(v.colours_used)) for ep, v in sorted(tr.all_validation_stats.items())}
    losses = [float(l.item()) for l in tr.losses[::max(1, len(tr.losses) // 50)]]
    res = {'episodes': episodes, 'math_mode': math_mode, 'wall_s': dt, 'episodes_per_sec': episodes / dt, 'env_steps': tr.t_step,
           'learn_steps': len(tr.losses), 'validation_mean_colours_by_episode': curve, 'dsatur_mean_colours': dsatur,
           'random_order_mean_colours': 8.5752, 'best_validation_mean_colours': min(curve.values()),
           'training_mean_colours_first_1000': float(np.mean(tr.training_stats.colours_used[:1000])),
           'training_mean_colours_last_1000': float(np.mean(tr.training_stats.colours_used[-1000:])),
           'loss_samples': losses,
           'what': 'scripts/train_full_schedule.py: the reference schedule on the device loop, 1000 synthetic training graphs, the reference\'s 100 validation graphs'}
    print(json.dumps(res))


if __name__ == '__main__':
    main()
