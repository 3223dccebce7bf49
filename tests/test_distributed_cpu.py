"""World-size-2 gloo test (CPU) of the N>1 learning rule through the function QLearner itself calls
(`trainer.reduce_mean_gradient`): every rank puts the SUMS of its shard -- flat gradient, loss, graph count -- into
one buffer, one all-reduce, and the result equals the single-process mean gradient on the whole minibatch, for an
UNEQUAL split (7 graphs over 2 ranks: 3 + 4; a per-rank 1/(n_local * world) scaling would be wrong here) and through
both normalisation routes (global size known to the caller / all-reduced with the gradient).  The per-shard
arithmetic here is the oracle's (the CUDA path cannot run on CPU; tests/test_multigpu_gpu.py runs the same check on
two real GPUs); also covered: the host-side shard and micro-batch helpers."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gcrl_testutil as util
from oracle import gn_oracle


def _loss_grads(params, obs, actions, targets, b_total):
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    x, ei, ea = gn_oracle.collate(obs)
    _, q = gn_oracle.gn_forward(pr, x, ei, ea)
    ns = [o[0].shape[0] for o in obs]
    off = np.concatenate([[0], np.cumsum(ns)[:-1]])
    qe = q.view(-1)[torch.as_tensor(off + actions)]
    loss = ((qe - torch.as_tensor(targets)) ** 2).sum() / b_total
    loss.backward()
    return float(loss.detach()), torch.cat([pr[k].grad.reshape(-1) for k in params])


def _make_batch():
    graphs = util.load_validation_graphs()
    rng = np.random.default_rng(0)
    gids = [1, 5, 9, 30, 47, 66, 12]
    obs = [gn_oracle.observation(graphs[g]['adj'], util.random_colours(graphs[g]['n'], graphs[g]['adj'], 0.4, rng),
                                 graphs[g]['edge_list']) for g in gids]
    actions = np.asarray([0, 3, 2, 7, 1, 4, 5])
    targets = rng.standard_normal(len(gids)).astype(np.float32)
    return obs, actions, targets


def _worker(rank, world, port, out_dir):
    from graph_colouring_with_rl_b200.trainer import shard_range
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    # a learner only some ranks run (the single-GPU training loop, bench.py's train_loop extra on rank 0) must not
    # pick up the default process group: it would enter all-reduces the other ranks never post
    from graph_colouring_with_rl_b200.trainer import QLearner
    assert QLearner(object()).world == world and QLearner(object(), local_only=True).world == 1
    from graph_colouring_with_rl_b200.trainer import reduce_mean_gradient
    params = util.load_parity_weights()
    obs, actions, targets = _make_batch()
    lo, hi = shard_range(len(obs), rank, world)
    assert (hi - lo) == (3 if rank == 0 else 4)
    out = {}
    for tag, total in (('count_allreduced', None), ('batch_total_given', len(obs))):
        # unscaled sums (b_total = 1) when the count rides in the buffer, else scaled by the known global size
        loss, grad = _loss_grads(params, obs[lo:hi], actions[lo:hi], targets[lo:hi], 1 if total is None else total)
        P = grad.numel()
        gbuf = torch.cat([grad, torch.tensor([loss, float(hi - lo)])])
        reduce_mean_gradient(gbuf, P, None, total)
        out[tag] = {'grad': gbuf[:P].clone(), 'loss': float(gbuf[P])}
    if rank == 0:
        torch.save(out, os.path.join(out_dir, 'reduced.pt'))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_single_process(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), 'reduced.pt'))
    params = util.load_parity_weights()
    obs, actions, targets = _make_batch()
    loss, grad = _loss_grads(params, obs, actions, targets, len(obs))
    scale = float(grad.abs().max())
    for tag in ('count_allreduced', 'batch_total_given'):
        assert abs(got[tag]['loss'] - loss) < 1e-5 * max(1.0, abs(loss)), tag
        assert float((got[tag]['grad'] - grad).abs().max()) < 5e-3 * scale, tag
    # what the round-1 rule (each rank scales by 1 / (n_local * world)) would have produced on this split: wrong
    l0, g0 = _loss_grads(params, obs[:3], actions[:3], targets[:3], 3 * 2)
    l1, g1 = _loss_grads(params, obs[3:], actions[3:], targets[3:], 4 * 2)
    assert float((g0 + g1 - grad).abs().max()) > 2e-2 * scale


def test_shard_and_micro_batch_helpers():
    from graph_colouring_with_rl_b200.trainer import shard_range
    for n, w in ((10000, 8), (7, 4), (3, 8), (0, 2)):
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    from graph_colouring_with_rl_b200.trainer import shard_host
    ns = np.asarray([3, 5, 2, 4])
    host = dict(adj=torch.arange(int((ns * ns).sum())), cur_colour=torch.arange(14), next_colour=torch.arange(14) + 100,
                action=torch.arange(4), reward=torch.arange(4.), done=torch.arange(4))
    n2, h2 = shard_host(ns, host, 1, 3)
    assert list(n2) == [5, 2] and h2['adj'].tolist() == list(range(9, 9 + 29)) and h2['cur_colour'].tolist() == list(range(3, 10))
    assert h2['next_colour'].tolist() == list(range(103, 110)) and h2['action'].tolist() == [1, 2]
