"""Launch list of DQN learn steps at the reference's own minibatch size (64 transitions, graphs of 15-50 vertices), for
   ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file out.csv python scripts/learn_step_small_launches.py
Also prints the wall-clock per learn step without the profiler's serialisation when run plainly."""
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
    tr.run(64, verbose=False)                       # fills the ring past the 640-transition gate
    for _ in range(5):
        tr.learn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    torch.cuda.profiler.start()
    n = 3 if os.environ.get('NV_NSIGHT_INJECTION_PORT_BASE') or os.environ.get('NCU') else 50
    for _ in range(n):
        tr.learn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    dt = time.perf_counter() - t0
print('%d learn steps, %.3f ms each (wall clock)' % (n, 1e3 * dt / n))
