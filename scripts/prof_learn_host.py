import os, sys, time, cProfile, pstats, io
sys.path.insert(0, '/root/repo')
import contextlib
import numpy as np
import torch
from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
from graph_colouring_with_rl_b200.dqn_main import DeviceTrainer
from graph_colouring_with_rl_b200.utils import erdos_renyi_adjacency
rng = np.random.default_rng(0)
graphs = [erdos_renyi_adjacency(int(n), 0.5, rng) for n in rng.integers(15, 51, size=128)]
torch.manual_seed(0)
agent = DQNAgentGN(2, 1, 0, math_mode='bf16x3')
with contextlib.redirect_stdout(sys.stderr):
    tr = DeviceTrainer(agent, graphs, None, n_envs=64, device_seed=0)
    tr.run(64, verbose=False)
    for _ in range(5):
        tr.learn()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    tr.learn()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28)
print(s.getvalue()[:6000])
