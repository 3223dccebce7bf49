"""Memory-safety and race checks of our own (compute-sanitizer is closed on this GPU pool, profiles/r02_sanitizer_unavailable.txt):
guard bands around every caller-owned buffer the library writes, and run-to-run determinism of the atomics-free forward
as a detector of shared-memory races in the tile reuse of the tensor-core kernels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GUARD = 4096
PATTERN = 0xA5


def er(n, p, rng):
    a = (rng.random((n, n)) < p).astype(np.uint8)
    a = np.triu(a, 1)
    return a | a.T


def guarded(nbytes):
    """uint8 buffer [GUARD | nbytes | GUARD] filled with PATTERN -> (whole, payload view)."""
    nbytes = (int(nbytes) + 255) // 256 * 256
    whole = torch.full((nbytes + 2 * GUARD,), PATTERN, dtype=torch.uint8, device=DEV)
    return whole, whole[GUARD:GUARD + nbytes]


def bands_intact(whole):
    return bool((whole[:GUARD] == PATTERN).all().item()) and bool((whole[-GUARD:] == PATTERN).all().item())


def make_batch(sizes, seed):
    from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
    rng = np.random.default_rng(seed)
    adjs = [er(n, 0.4, rng) for n in sizes]
    if len(adjs) > 3:
        adjs[3][:] = 0                                  # isolated vertices
    env = BatchedColouringEnv(adjs, device=DEV)
    env.step(env.first_vertices())
    return env


@pytest.mark.parametrize('math_mode', [0, 2, 1])
@pytest.mark.parametrize('sizes', [[2, 3, 5, 4, 2, 9, 17, 3, 33, 2], [200, 7], [64, 129, 31]])
def test_guard_bands_and_forward_determinism(sizes, math_mode):
    from graph_colouring_with_rl_b200 import _lib, ops
    from oracle import gn_oracle
    import gcrl_testutil as util
    lib = _lib.load()
    env = make_batch(sizes, 3)
    packed, eattr = env.packed
    N, E = packed.n_vertices, packed.n_edges
    params = util.load_parity_weights()
    flat = util.flatten(params).to(DEV)
    stash_w, stash = guarded(lib.gcrl_gn_stash_bytes(N, E))
    ws_w, ws = guarded(lib.gcrl_gn_forward_workspace_bytes(N, E, 1))
    outs = []
    for rep in range(5):
        h, q, st = ops.gn_forward(flat, (2, 1, 0), packed, env.x, eattr, with_stash=True, math_mode=math_mode, stash=stash, ws=ws)
        assert st.data_ptr() == stash.data_ptr()
        outs.append((h.clone(), q.clone(), stash.clone()))
    torch.cuda.synchronize()
    assert bands_intact(stash_w) and bands_intact(ws_w), 'forward wrote outside its stash / workspace'
    for h, q, s in outs[1:]:
        assert torch.equal(h, outs[0][0]) and torch.equal(q, outs[0][1]), 'forward is not run-to-run deterministic'
        # the stash holds padding between its arrays (never written): compare what the library wrote through the same mask
        assert torch.equal(s, outs[0][2]), 'activation stash differs between runs'
    # inference flavour (no stash, e ping-pong in the workspace)
    ws2_w, ws2 = guarded(lib.gcrl_gn_forward_workspace_bytes(N, E, 0))
    q0 = None
    for rep in range(3):
        _, q, _ = ops.gn_forward(flat, (2, 1, 0), packed, env.x, eattr, math_mode=math_mode, ws=ws2)
        q0 = q.clone() if q0 is None else q0
        assert torch.equal(q, q0)
    assert bands_intact(ws2_w)
    tol = {0: 2e-5, 2: 3e-4, 1: 0.15}[math_mode]
    assert float((q0 - outs[0][1]).abs().max()) <= tol            # stash / no-stash flavours agree
    # backward: gradient buffer and workspace guarded; repeat runs agree (float atomics: order differs) and are finite
    grad_w, grad_b = guarded(flat.numel() * 4)
    grad = grad_b.view(torch.float32)[:flat.numel()]
    wsb_w, wsb = guarded(lib.gcrl_gn_backward_workspace_bytes(N, E))
    torch.manual_seed(0)
    d_q = torch.randn(N, device=DEV)
    gs = []
    for rep in range(3):
        ops.gn_backward(flat, (2, 1, 0), packed, env.x, eattr, stash, d_q=d_q, grad=grad, accumulate=False, math_mode=math_mode, ws=wsb)
        gs.append(grad.clone())
    torch.cuda.synchronize()
    assert bands_intact(grad_w) and bands_intact(wsb_w) and bands_intact(stash_w), 'backward wrote outside its buffers'
    scale = float(gs[0].abs().max())
    assert np.isfinite(scale) and scale > 0
    for g in gs[1:]:
        assert float((g - gs[0]).abs().max()) <= 1e-4 * scale        # float atomics: the summation order differs run to run


def test_rollouts_identical_eager_graphed_and_without_pdl():
    """The CUDA-graph replay of the rollout step, and programmatic dependent launch, change scheduling only: action
    sequences and colours are identical to the eager run."""
    import os
    import subprocess
    import sys
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    from graph_colouring_with_rl_b200.rollouts import greedy_rollouts
    rng = np.random.default_rng(9)
    adjs = [er(int(n), 0.5, rng) for n in rng.integers(8, 40, size=48)]
    torch.manual_seed(0)
    agent = DQNAgentGN(2, 1, 0, test_mode=True, math_mode='bf16x3')
    with torch.no_grad():
        agent.gn_behaviour.flat_params().mul_(1.5)
    a = greedy_rollouts(adjs, agent.gn_behaviour, record=True, graph=False)
    b = greedy_rollouts(adjs, agent.gn_behaviour, record=True, graph=True)
    c = greedy_rollouts(None, agent.gn_behaviour, record=True, graph=True, env=b['env'])       # cached graph, second sweep
    for r in (b, c):
        assert r['steps'] == a['steps'] and torch.equal(r['colours_used'], a['colours_used'])
        assert all(torch.equal(x, y) for x, y in zip(r['actions'], a['actions']))
    if os.environ.get('GCRL_PDL') != '0':
        # the same sweep in a child process with PDL off
        code = ('import sys, numpy as np, torch; sys.path[:0] = [%r, %r]\n'
                'from test_guards_gpu import er\n'
                'from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN\n'
                'from graph_colouring_with_rl_b200.rollouts import greedy_rollouts\n'
                'rng = np.random.default_rng(9); adjs = [er(int(n), 0.5, rng) for n in rng.integers(8, 40, size=48)]\n'
                'torch.manual_seed(0); agent = DQNAgentGN(2, 1, 0, test_mode=True, math_mode="bf16x3")\n'
                'agent.gn_behaviour.flat_params().data.mul_(1.5)\n'
                'r = greedy_rollouts(adjs, agent.gn_behaviour, record=True, graph=False)\n'
                'print("RESULT", r["colours_used"].cpu().tolist(), torch.stack(r["actions"]).cpu().tolist())\n'
                % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__))))
        out = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, GCRL_PDL='0'), capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        line = [l for l in out.stdout.splitlines() if l.startswith('RESULT')][0]
        want = 'RESULT %s %s' % (a['colours_used'].cpu().tolist(), torch.stack(a['actions']).cpu().tolist())
        assert line == want


def test_env_step_greedy_equals_argmax_then_step():
    """gcrl_env_step_greedy (masked first-argmax fused into the environment step) == gcrl_score_argmax followed by
    gcrl_env_step, bit for bit: colours, colour counts, rewards, done flags, observations and the actions taken, with
    exact Q ties (first maximum = lowest vertex index, dqn_agent.py:131-134) and finished graphs (action -1)."""
    from graph_colouring_with_rl_b200 import ops
    from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
    rng = np.random.default_rng(4)
    adjs = [er(int(n), 0.35, rng) for n in rng.integers(1, 70, size=97)] + [er(300, 0.1, rng)]
    e1 = BatchedColouringEnv(adjs, device=DEV)
    e2 = BatchedColouringEnv(adjs, device=DEV)
    gen = torch.Generator(device=DEV).manual_seed(1)
    for step in range(75):
        q = torch.randint(0, 4, (e1.n_vertices,), generator=gen, device=DEV).float()       # many exact ties
        a1 = ops.score_argmax(q, e1.colour, e1.node_off)
        e1.step(a1)
        a2 = e2.step_greedy(q)
        assert torch.equal(a1, a2), step
        for k in ('colour', 'max_colour', 'reward', 'done', 'x'):
            assert torch.equal(getattr(e1, k), getattr(e2, k)), (step, k)
    assert bool((a2[:97] == -1).all())          # the small graphs finished long ago
