"""How many graphs are still running at each step of a greedy rollout sweep (sizing the benefit of compaction)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
from graph_colouring_with_rl_b200 import ops
Gr, nr = 2000, 50
adjs = bench.er_graphs(Gr, nr, 0.5, seed=0)
torch.manual_seed(0)
agent = DQNAgentGN(2, 1, 0, math_mode='bf16x3')
net = agent.gn_behaviour
env = BatchedColouringEnv(list(adjs), device=net.device)
packed, eattr = env.packed
flat = net.flat_params()
a = env.first_vertices(); env.step(a)
alive = []
for s in range(1, nr):
    alive.append(int((env.done == 0).sum().item()))
    _, q, _ = ops.gn_forward(flat, net.dims, packed, env.x, eattr, math_mode=net.math_mode)
    a = ops.score_argmax(q, env.colour, env.node_off)
    env.step(a)
print('alive per step', alive)
print('sum alive / (G*steps) = %.3f' % (sum(alive) / (Gr * len(alive))))
