"""Parity at the BENCHMARKED shape: `QLearner.fwd_bwd` -- the call bench.py times -- on Erdos-Renyi graphs of 130,
200 and 257 vertices (segments of 129 / 199 / 256 edges span two to three 128-row tiles: cross-tile dPi column sums,
eight-row dPj reduce boxes, arg-term routing across tile boundaries), split into four micro-batches of unequal size so
that the gradient is accumulated over DIFFERENT slices (`accumulate=True`), in all three math modes, against the CPU
oracle (oracle/gn_oracle.py; reference: dqn_agent.py:174-220): loss, Q_expected, Q_target and every gradient tensor.

The oracle is evaluated graph by graph (graphs of a minibatch are independent, dqn_agent.py:55-56 builds a
block-diagonal batch) and summed with the 1/B weights of the batch mean, once per test session for all modes.
"""
import numpy as np
import pytest
import torch

import gcrl_testutil as util
from oracle import gn_oracle

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
NS = [130, 130, 200, 257, 200, 257, 130]
MAX_EDGES = 100000          # micro-batches: [130,130,200] [257] [200] [257,130] (73 340 / 65 792 / 39 800 / 82 562 edges)
_cache = {}


def er(n, p, rng):
    a = (rng.random((n, n)) < p).astype(np.uint8)
    a = np.triu(a, 1)
    return a | a.T


def transitions():
    """Seeded mid-episode transitions from the device environment (itself bit-exact against the reference's
    traces, tests/test_kernels_gpu.py) + the oracle's gradients for them."""
    if 'host' in _cache:
        return _cache
    from graph_colouring_with_rl_b200.trainer import synthetic_transitions
    rng = np.random.default_rng(20)
    adjs = [er(n, 0.5, rng) for n in NS]
    ns, host = synthetic_transitions(adjs, 0.5, seed=3, device=DEV)
    assert list(ns) == NS and int(host['action'].min()) >= 0
    params = util.load_parity_weights()
    B = len(NS)
    node_off = np.concatenate([[0], np.cumsum(NS)])
    cur, nxt = host['cur_colour'].numpy(), host['next_colour'].numpy()
    act, rew, done = host['action'].numpy(), host['reward'].numpy(), host['done'].numpy()

    def evaluate(dtype, permute):
        gen = torch.Generator().manual_seed(5)
        grads = {k: torch.zeros_like(v, dtype=torch.float64) for k, v in params.items()}
        loss, q_exp, q_tgt = 0.0, [], []
        for g, n in enumerate(NS):
            v0, v1 = node_off[g], node_off[g + 1]
            x, ei, ea = gn_oracle.observation(adjs[g], cur[v0:v1])
            xn, _, _ = gn_oracle.observation(adjs[g], nxt[v0:v1])
            if permute:
                p = torch.randperm(ei.shape[1], generator=gen)
                ei, ea = ei[:, p], ea[p]
            with torch.no_grad():                                   # target network = same weights here
                _, qn = gn_oracle.gn_forward({k: v.to(dtype) for k, v in params.items()}, xn.to(dtype), ei, ea.to(dtype))
            tgt = gn_oracle.td_targets(qn.float().numpy(), (nxt[v0:v1] >= 0).astype(np.int64), [n], rew[g:g + 1],
                                       done[g:g + 1], 1.0)
            pr = {k: v.clone().to(dtype).requires_grad_(True) for k, v in params.items()}
            _, q = gn_oracle.gn_forward(pr, x.to(dtype), ei, ea.to(dtype))
            qe = q.view(-1)[int(act[g])]
            lg = (qe - float(tgt[0])) ** 2 / B
            lg.backward()
            for k in params:
                grads[k] += pr[k].grad.double()
            loss += float(lg.detach())
            q_exp.append(float(qe.detach()))
            q_tgt.append(float(tgt[0]))
        flat = torch.cat([grads[k].reshape(-1) for k in params]).numpy()
        return dict(loss=loss, q_exp=np.asarray(q_exp), q_tgt=np.asarray(q_tgt), grad=flat)
    _cache.update(host=host, ns=ns, params=params, truth=evaluate(torch.float64, False),
                  refs=[evaluate(torch.float32, False), evaluate(torch.float32, True)])
    return _cache


def make_learner(math_mode, params):
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    from graph_colouring_with_rl_b200.trainer import QLearner
    torch.manual_seed(0)
    agent = DQNAgentGN(2, 1, 0, math_mode=math_mode)
    flat = util.flatten(params).to(DEV)
    agent.gn_behaviour.flat_params().copy_(flat)
    agent.gn_target.flat_params().copy_(flat)
    learner = QLearner(agent, max_edges_per_micro_batch=MAX_EDGES, local_only=True)
    learner.keep_q = True
    return agent, learner


@pytest.mark.parametrize('math_mode', ['fp32', 'bf16x3', 'bf16'])
def test_qlearner_fwd_bwd_at_the_benchmarked_shape_vs_oracle(math_mode):
    from test_kernels_gpu import grads_close
    from graph_colouring_with_rl_b200.trainer import TransitionBatch
    c = transitions()
    params, truth, refs = c['params'], c['truth'], c['refs']
    agent, learner = make_learner(math_mode, params)
    batch = TransitionBatch.from_host(c['ns'], c['host'], DEV)
    ranges = batch.micro_ranges(MAX_EDGES)
    assert len(ranges) >= 3 and len(set(g1 - g0 for g0, g1 in ranges)) >= 2, ranges      # unequal micro-batches
    q_next = learner.target_q(batch)
    loss = learner.fwd_bwd(batch, q_next)
    got = agent.gn_behaviour._flat_grad.cpu().numpy()
    q_exp, q_tgt = [t.cpu().numpy() for t in learner.last_q]
    assert np.all(np.isfinite(got))
    # tolerances: |Q| ~ 1 on the trained weights.  fp32: the parity path; bf16x3: ~16 mantissa bits per operand;
    # bf16: a throughput mode (8 bits), stated separately
    q_tol = {'fp32': 5e-5, 'bf16x3': 3e-4, 'bf16': 8e-2}[math_mode]
    np.testing.assert_allclose(q_exp, truth['q_exp'], atol=q_tol, rtol=0)
    np.testing.assert_allclose(q_tgt, truth['q_tgt'], atol=q_tol, rtol=0)
    l_tol = {'fp32': 1e-4, 'bf16x3': 1e-3, 'bf16': 0.25}[math_mode]
    assert abs(float(loss.item()) - truth['loss']) <= l_tol * max(1.0, abs(truth['loss'])), (float(loss.item()), truth['loss'])
    r32 = [r['grad'] for r in refs]
    if math_mode == 'fp32':
        grads_close(got, truth['grad'], r32, params)
    elif math_mode == 'bf16x3':
        grads_close(got, truth['grad'], r32, params, rtol=2e-3, noise_amp=128.0)
    else:
        g, t = got.astype(np.float64), truth['grad']
        cos = float(g @ t / (np.linalg.norm(g) * np.linalg.norm(t)))
        rel = float(np.linalg.norm(g - t) / np.linalg.norm(t))
        print('bf16: cosine %.4f, relative L2 error %.3f' % (cos, rel))
        assert cos >= 0.97 and rel <= 0.25, (cos, rel)
    # the same minibatch in ONE micro-batch gives the same gradient (accumulation over slices is exact up to fp32 order)
    _, one = make_learner(math_mode, params)
    one.max_edges = 10 ** 9
    assert len(batch.micro_ranges(one.max_edges)) == 1
    loss1 = one.fwd_bwd(batch, q_next)
    got1 = one.agent.gn_behaviour._flat_grad.cpu().numpy()
    scale = np.abs(truth['grad']).max()
    assert np.abs(got1 - got).max() <= {'fp32': 2e-4, 'bf16x3': 5e-4, 'bf16': 5e-2}[math_mode] * scale
    assert abs(float(loss1.item()) - float(loss.item())) <= 1e-5 * max(1.0, abs(truth['loss']))


def test_bad_action_is_flagged_not_dereferenced():
    """a sentinel action (-1) or one past the graph must not read / write out of bounds."""
    from graph_colouring_with_rl_b200 import ops
    node_off = torch.tensor([0, 4, 9], dtype=torch.int32, device=DEV)
    q = torch.arange(9, dtype=torch.float32, device=DEV)
    col = torch.full((9,), -1, dtype=torch.int32, device=DEV)
    rew = torch.zeros(2, device=DEV)
    done = torch.zeros(2, dtype=torch.int32, device=DEV)
    for bad in (-1, 5):
        act = torch.tensor([2, bad], dtype=torch.int32, device=DEV)
        err = torch.zeros(1, dtype=torch.int32, device=DEV)
        loss, q_exp, q_tgt, d_q = ops.td_target_loss(q, q, col, node_off, act, rew, done, 1.0, err=err)
        assert int(err.item()) == 2
        assert torch.isnan(q_exp[1]) and float(q_exp[0]) == 2.0
        assert float(d_q[4:].abs().sum()) == 0.0 and float(d_q[2]) != 0.0
        assert np.isfinite(float(loss.item()))
