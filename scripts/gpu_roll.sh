cd /root/repo
timeout 600 python -m pytest tests -q -x --timeout 300 -m gpu 2>&1 | tail -1
python - <<'PY'
import sys, os, time, torch
sys.path.insert(0, '/root/repo')
import bench
from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
from graph_colouring_with_rl_b200.rollouts import greedy_rollouts
Gr, nr = 10000, 50
adjs = bench.er_graphs(Gr, nr, 0.5, seed=0)
torch.manual_seed(0)
agent = DQNAgentGN(2, 1, 0, math_mode='bf16x3')
env = BatchedColouringEnv(list(adjs), device=agent.gn_behaviour.device)
greedy_rollouts(None, agent.gn_behaviour, env=env)
for _ in range(3):
    torch.cuda.synchronize(); t=time.perf_counter()
    res = greedy_rollouts(None, agent.gn_behaviour, env=env)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('rollout sweep %.1f ms -> %.0f graphs/s' % (dt*1e3, Gr/dt))
PY
