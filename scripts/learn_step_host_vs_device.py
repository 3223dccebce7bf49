"""Where does a 64-transition learn step spend its time: host issue time (Python + ctypes + launches, no sync) vs
device time (CUDA events), per phase."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
mem, learner = tr.memory, tr.learner
torch.cuda.synchronize()
N = 50
phases = ['sample+gather', 'target_q', 'fwd_bwd', 'adam']
host = {p: 0.0 for p in phases}
dev = {p: 0.0 for p in phases}
for it in range(N):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    evs[0].record()
    t0 = time.perf_counter()
    batch = mem.gather(mem.sample_indices(agent.batch_size))
    t1 = time.perf_counter(); evs[1].record()
    q_next = learner.target_q(batch)
    t2 = time.perf_counter(); evs[2].record()
    loss = learner.fwd_bwd(batch, q_next)
    t3 = time.perf_counter(); evs[3].record()
    agent.gn_behaviour.optimizer.step(target_flat=agent.gn_target.flat_params(), tau=agent.tau, grad=agent.gn_behaviour._flat_grad)
    t4 = time.perf_counter(); evs[4].record()
    torch.cuda.synchronize()
    for p, a, b in zip(phases, (t0, t1, t2, t3), (t1, t2, t3, t4)):
        host[p] += (b - a) * 1e3
    for i, p in enumerate(phases):
        dev[p] += evs[i].elapsed_time(evs[i + 1])
print('phase            host issue ms   device ms (events, includes waiting for the host)')
for p in phases:
    print('%-16s %8.3f        %8.3f' % (p, host[p] / N, dev[p] / N))
print('total            %8.3f        %8.3f' % (sum(host.values()) / N, sum(dev.values()) / N))
# back-to-back without per-step sync
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(N):
    tr.learn()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print('back to back: host issue %.3f ms/step, wall %.3f ms/step' % (1e3 * t_issue / N, 1e3 * t_all / N))
