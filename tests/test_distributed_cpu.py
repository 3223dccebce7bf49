"""World-size-2 gloo test (CPU) of the N>1 learning rule: each rank scales its
minibatch shard's loss by 1/B_global, the flat gradients are summed with one all-reduce, and the
result equals the single-process gradient on the whole minibatch.  The arithmetic here is the
oracle's (the CUDA path cannot run on CPU); what is under test is the sharding/all-reduce contract
QLearner implements, plus the host-side shard and micro-batch helpers."""
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
    gids = [1, 5, 9, 30, 47, 66]
    obs = [gn_oracle.observation(graphs[g]['adj'], util.random_colours(graphs[g]['n'], graphs[g]['adj'], 0.4, rng),
                                 graphs[g]['edge_list']) for g in gids]
    actions = np.asarray([0, 3, 2, 7, 1, 4])
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
    params = util.load_parity_weights()
    obs, actions, targets = _make_batch()
    lo, hi = shard_range(len(obs), rank, world)
    loss, grad = _loss_grads(params, obs[lo:hi], actions[lo:hi], targets[lo:hi], len(obs))
    loss_t = torch.tensor([loss])
    dist.all_reduce(grad)
    dist.all_reduce(loss_t)
    if rank == 0:
        torch.save({'grad': grad, 'loss': float(loss_t)}, os.path.join(out_dir, 'reduced.pt'))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_single_process(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), 'reduced.pt'))
    params = util.load_parity_weights()
    obs, actions, targets = _make_batch()
    loss, grad = _loss_grads(params, obs, actions, targets, len(obs))
    assert abs(got['loss'] - loss) < 1e-5 * max(1.0, abs(loss))
    scale = float(grad.abs().max())
    assert float((got['grad'] - grad).abs().max()) < 5e-3 * scale


def test_shard_and_micro_batch_helpers():
    from graph_colouring_with_rl_b200.trainer import shard_range
    for n, w in ((10000, 8), (7, 4), (3, 8), (0, 2)):
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
