cd /root/repo
timeout 900 python -m pytest tests -q -x --timeout 300 -m gpu 2>&1 | tail -12
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
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter()
    res = greedy_rollouts(None, agent.gn_behaviour, env=env)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('rollout sweep %.1f ms -> %.0f graphs/s  mean colours %.4f' % (dt*1e3, Gr/dt, res['colours_used'].float().mean().item()))
PY
timeout 300 python bench.py --graphs 512 --micro-graphs 256 --skip-extras --steps 3 --warmup 1 2>gpurun_out/bench_ws.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench', round(d['value']), round(d['ms_per_step'],1))"
