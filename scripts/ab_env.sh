#!/bin/bash
# A/B of an environment switch of the library on ONE box, alternating runs:
#   bash scripts/ab_env.sh GCRL_FIRST_DIRECT=0 [tag]
# prints the fwd+bwd headline (short run) with the per-kernel averages, and one rollout sweep (10 000 x 50-vertex graphs).
SW=$1; TAG=${2:-alt}
run() { env $1 timeout 300 python bench.py --skip-extras --steps 8 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$2 fwd+bwd', round(d['value']), d['clocks']['sm_mhz'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"; }
roll() { env $1 timeout 300 python - <<'EOF'
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
from graph_colouring_with_rl_b200.rollouts import greedy_rollouts
torch.manual_seed(0)
agent = DQNAgentGN(2, 1, 0, test_mode=True, math_mode='bf16x3')
env = BatchedColouringEnv(list(bench.er_graphs(10000, 50, 0.5, seed=0)), device=agent.gn_behaviour.device)
greedy_rollouts(None, agent.gn_behaviour, env=env)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(2):
    a.record(); r = greedy_rollouts(None, agent.gn_behaviour, env=env); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print('rollouts graphs/s', round(10000 / best * 1e3), 'ms/sweep', round(best, 1), 'steps', r['steps'], 'colours', float(r['colours_used'].float().mean()))
EOF
}
run "X=1" default; run "$SW" $TAG; run "X=1" default; run "$SW" $TAG
echo default; roll "X=1"; echo $TAG; roll "$SW"
