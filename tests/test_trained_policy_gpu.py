"""Colouring parity with a FULLY trained policy.  tests/golden/trained25k_GN is the behaviour network after the reference's
whole training schedule (25 000 episodes) run by this repository's device training loop on a B200
(scripts/train_full_schedule.py; the reference's own trained_policies/learned_parameters_GN is not in its checkout);
tests/golden/rollouts_trained25k.npz holds what the UNMODIFIED reference does with those weights on its 100 validation graphs on
the CPU (oracle/make_golden.py rollouts_trained25k: vertex sequences, colours used, the top-2 Q gap of every decision).

The device must colour identically wherever the reference's Q gap exceeds the mode's tolerance (north_star), use the same
number of colours on those graphs, and land on the reference's mean colour count (7.45; DSATUR 7.42, random order 8.58)."""
import os

import numpy as np
import pytest
import torch

import gcrl_testutil as util

pytestmark = pytest.mark.gpu
GRAPHS = util.load_validation_graphs()


# tol: twice the mode's stated Q tolerance on trained weights (5e-5 / 3e-4); min_full: graphs ALL of whose decisions have a
# larger gap in the reference's run (73 / 30 of the 100)
@pytest.mark.parametrize('math_mode,tol,min_full', [('fp32', 1e-4, 70), ('bf16x3', 6e-4, 25)])
def test_trained_policy_colours_like_the_reference(math_mode, tol, min_full):
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    from graph_colouring_with_rl_b200.rollouts import greedy_rollouts, sequences_from_actions
    d = np.load(os.path.join(util.GOLD, 'rollouts_trained25k.npz'))
    torch.manual_seed(0)
    agent = DQNAgentGN(2, 1, 0, test_mode=True, math_mode=math_mode)
    agent.load_models(os.path.join(util.GOLD, 'trained25k'))
    res = greedy_rollouts([dict(adj=g['adj'], name=g['name']) for g in GRAPHS], agent.gn_behaviour, record=True)
    seqs = sequences_from_actions(res['actions'])
    used = res['colours_used'].cpu().numpy()
    ref_seqs = util.split_by(d['seq'], d['seq_len'])
    gaps = util.split_by(d['gaps'], d['gaps_len'])
    full = decisions = same_seq = same_used = 0
    for i in range(100):
        ref = [int(v) for v in ref_seqs[i]]
        same_seq += seqs[i] == ref
        same_used += int(used[i] == d['colours_used'][i])
        if gaps[i].size == 0 or gaps[i].min() > tol:
            assert seqs[i] == ref, GRAPHS[i]['name']
            assert used[i] == d['colours_used'][i]
            full += 1
            decisions += len(ref)
        else:                                   # identical up to the first decision the reference itself made on a near-tie
            k = int(np.argmax(gaps[i] <= tol)) + 1          # (+1: the first vertex comes from the neighbour-count rule)
            assert seqs[i][:k] == ref[:k], GRAPHS[i]['name']
            decisions += k
    print('%s: %d graphs must match to the end (every gap > %g) and do, %d decisions compared; regardless of gaps %d of 100 vertex '
          'sequences and %d of 100 colour counts are identical; mean colours %.2f (reference %.2f)'
          % (math_mode, full, tol, decisions, same_seq, same_used, used.mean(), d['colours_used'].mean()))
    assert full >= min_full
    assert abs(float(used.mean()) - float(d['colours_used'].mean())) <= 0.1
    col = res['env'].colour.cpu().numpy()
    off = res['env'].node_off.cpu().numpy()
    for i, g in enumerate(GRAPHS):
        c = col[off[i]:off[i + 1]]
        assert c.min() >= 0 and not np.any(g['adj'].astype(bool) & (c[:, None] == c[None, :]))
