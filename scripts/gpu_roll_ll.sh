cd /root/repo
cat > /tmp/roll_ll.py <<'PY'
import sys, os, time, torch
sys.path.insert(0, '/root/repo')
import bench
from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
from graph_colouring_with_rl_b200 import ops
Gr, nr = 10000, 50
adjs = bench.er_graphs(Gr, nr, 0.5, seed=0)
torch.manual_seed(0)
agent = DQNAgentGN(2, 1, 0, math_mode='bf16x3')
net = agent.gn_behaviour
env = BatchedColouringEnv(list(adjs), device=net.device)
packed, eattr = env.packed
flat = net.flat_params()
a = env.first_vertices(); env.step(a)
for s in range(6):
    if s == 3: torch.cuda.synchronize(); torch.cuda.profiler.start()
    _, q, _ = ops.gn_forward(flat, net.dims, packed, env.x, eattr, math_mode=net.math_mode)
    a = ops.score_argmax(q, env.colour, env.node_off)
    env.step(a)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
PY
python /tmp/roll_ll.py && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_roll.csv python /tmp/roll_ll.py > gpurun_out/ncu_roll.log 2>&1
