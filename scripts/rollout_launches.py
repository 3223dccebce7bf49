"""Launch list of greedy-rollout steps (10 000 x 50-vertex graphs) for `ncu --metrics gpu__time_duration.sum`:
   ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file out.csv python scripts/rollout_launches.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graph_colouring_with_rl_b200 import ops
from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
from graph_colouring_with_rl_b200.utils import erdos_renyi_adjacency

rng = np.random.default_rng(0)
graphs = [erdos_renyi_adjacency(50, 0.5, rng) for _ in range(int(os.environ.get('ROLL_GRAPHS', '10000')))]
torch.manual_seed(0)
agent = DQNAgentGN(2, 1, 0, test_mode=True, math_mode='bf16x3')
net = agent.gn_behaviour
env = BatchedColouringEnv(graphs, device=net.device)
packed, eattr = env.packed
flat = net.flat_params()
env.step(env.first_vertices())
def step():
    _, q, _ = ops.gn_forward(flat, net.dims, packed, env.x, eattr, math_mode=net.math_mode)
    env.step_greedy(q)
for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
