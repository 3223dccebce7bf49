"""Worker of tests/test_multigpu_gpu.py (run under torchrun, one rank per GPU): the k-rank learning path of
`QLearner` against the 1-rank path on the union minibatch.

Checks, on every rank unless noted:
  * ranks that seeded differently hold rank 0's parameters after QLearner's broadcast;
  * each rank's shard of the graph layout (seg_off / src / dst / perm from gcrl_pack_dense) is BIT-EXACT the
    corresponding slice of the union layout, shifted by the shard's vertex / edge offsets;
  * the all-reduced flat gradient and loss equal the single-GPU gradient / loss of the union minibatch -- for an
    UNEQUAL split (7 graphs over 2 ranks: 3 + 4) through both normalisation routes (batch_total given / all-reduced
    graph count) -- to fp32 summation-order tolerance;
  * after a full learn step the parameters and the target parameters are bit-identical on all ranks and equal to
    the single-GPU step within Adam's sensitivity.
Writes <out>/mgpu_rank0.json with the measured differences (rank 0)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]


def er(n, p, rng):
    a = (rng.random((n, n)) < p).astype(np.uint8)
    a = np.triu(a, 1)
    return a | a.T


def train_main():
    """Multi-GPU DeviceTrainer (dqn_main.py:158-199 sharded by graph): after 3 waves of 16 episodes on 2 ranks the
    replicas hold bit-identical parameters, both ranks took the same number of learn steps, the learn cadence matches
    the global environment-step count, and the gathered statistics cover every episode."""
    out_dir = sys.argv[1]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    os.environ['GCRL_DEVICE'] = 'cuda:%d' % local
    dist.init_process_group('nccl', device_id=dev)
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    from graph_colouring_with_rl_b200.dqn_main import DeviceTrainer
    rng = np.random.default_rng(5)
    graphs = [er(int(n), 0.5, rng) for n in rng.integers(12, 31, size=24)]
    val = [dict(adj=er(int(n), 0.4, rng), name='v%d' % i) for i, n in enumerate(rng.integers(10, 25, size=7))]
    torch.manual_seed(50 + rank)                      # different seeds: the learner's broadcast makes the replicas equal
    np.random.seed(3 + rank)
    agent = DQNAgentGN(2, 1, 0, math_mode='bf16x3')
    agent.batch_size = 16                             # gate = 160 stored transitions: reached in the second wave
    tr = DeviceTrainer(agent, graphs, val, n_envs=16, device_seed=1, validate_every=32, distributed=True)
    assert tr.world == world and tr.rank == rank
    tr.run(48, verbose=False)
    assert tr.episodes_done == 48 and len(tr.training_stats.colours_used) == 48
    n_learn = torch.tensor([len(tr.losses), tr.t_step], device=dev)
    both = [torch.empty_like(n_learn) for _ in range(world)]
    dist.all_gather(both, n_learn)
    assert all(torch.equal(b, both[0]) for b in both), both
    assert len(tr.losses) > 0, 'the gate was never passed'
    assert all(np.isfinite(float(l.item())) for l in tr.losses)
    for net in (agent.gn_behaviour, agent.gn_target):
        p = net.flat_params()
        parts = [torch.empty_like(p) for _ in range(world)]
        dist.all_gather(parts, p.contiguous())
        assert all(torch.equal(q, parts[0]) for q in parts), 'replicas diverged'
    # validation ran at episodes 0, 32 on every rank with the same gathered result
    assert sorted(tr.all_validation_stats.keys()) == [0, 32]
    assert len(tr.all_validation_stats[32].colours_used) == len(val)
    if rank == 0:
        with open(os.path.join(out_dir, 'mgpu_train_rank0.json'), 'w') as f:
            json.dump({'world': world, 'episodes': tr.episodes_done, 'env_steps': tr.t_step, 'learn_steps': len(tr.losses),
                       'last_loss': float(tr.losses[-1].item()), 'mean_colours_last16': float(np.mean(tr.training_stats.colours_used[-16:])),
                       'validation_colours_ep32': float(np.mean(tr.all_validation_stats[32].colours_used))}, f)
    dist.barrier()
    dist.destroy_process_group()


def main():
    out_dir, math_mode = sys.argv[1], sys.argv[2]
    if math_mode == 'train':
        return train_main()
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    os.environ['GCRL_DEVICE'] = 'cuda:%d' % local
    dist.init_process_group('nccl', device_id=dev)
    import gcrl_testutil as util
    from graph_colouring_with_rl_b200 import ops
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    from graph_colouring_with_rl_b200.trainer import (QLearner, TransitionBatch, shard_batch, shard_range,
                                                      synthetic_transitions)
    ns_all = [40, 130, 64, 200, 25, 257, 90]
    rng = np.random.default_rng(11)
    adjs = [er(n, 0.5, rng) for n in ns_all]
    ns, host = synthetic_transitions(adjs, 0.5, seed=4, device=dev)          # identical on every rank (seeded)
    params = util.load_parity_weights()
    flat0 = util.flatten(params)

    def fresh_agent(seed, load):
        torch.manual_seed(seed)
        agent = DQNAgentGN(2, 1, 0, math_mode=math_mode)
        if load:
            agent.gn_behaviour.flat_params().copy_(flat0.to(dev))
            agent.gn_target.flat_params().copy_(flat0.to(dev))
        return agent

    report = {}
    # ---- parameter broadcast: rank r seeds with r, rank 0 holds the parity weights ------------------------------
    agent = fresh_agent(100 + rank, load=(rank == 0))
    learner = QLearner(agent, max_edges_per_micro_batch=70000)
    assert learner.world == world
    got = agent.gn_behaviour.flat_params().cpu()
    assert torch.equal(got, flat0), 'rank %d does not hold rank 0 parameters after the broadcast' % rank
    assert torch.equal(agent.gn_target.flat_params().cpu(), flat0)

    # ---- shard layout bit-exact vs the union layout ------------------------------------------------------------------
    union = TransitionBatch.from_host(ns, host, dev)
    lo, hi = shard_range(len(ns_all), rank, world)
    mine = shard_batch(ns, host, lo, hi, dev)
    assert mine.n_graphs == hi - lo
    pu, eu = ops.pack_dense(union.adj, union.adj_off, union.node_off, union.n_vertices, union.n_edges)
    pm, em = ops.pack_dense(mine.adj, mine.adj_off, mine.node_off, mine.n_vertices, mine.n_edges)
    v0, e0 = int(union.node_off_h[lo]), int(union.edge_off_h[lo])
    v1, e1 = int(union.node_off_h[hi]), int(union.edge_off_h[hi])
    assert torch.equal(pm.seg_off, pu.seg_off[v0:v1 + 1] - e0)
    assert torch.equal(pm.src, pu.src[e0:e1] - v0) and torch.equal(pm.dst, pu.dst[e0:e1] - v0)
    assert torch.equal(pm.perm, pu.perm[e0:e1] - e0) and torch.equal(em, eu[e0:e1])

    # ---- single-GPU reference on the union (every rank computes it locally, no communication) ------------------------
    solo_agent = fresh_agent(0, load=True)
    solo = QLearner(solo_agent, max_edges_per_micro_batch=70000, local_only=True)
    q_next_u = solo.target_q(union)
    loss_u = float(solo.fwd_bwd(union, q_next_u).item())
    grad_u = solo_agent.gn_behaviour._flat_grad.clone()
    scale = float(grad_u.abs().max())

    # ---- k-rank gradient, both normalisation routes ---------------------------------------------------------------------
    q_next_m = learner.target_q(mine)
    assert float((q_next_m - q_next_u[v0:v1]).abs().max()) <= 1e-4, 'target Q of a shard differs from the union forward'
    for tag, total in (('count_allreduced', None), ('batch_total_given', len(ns_all))):
        loss_k = float(learner.fwd_bwd(mine, q_next_m, batch_total=total).item())
        grad_k = agent.gn_behaviour._flat_grad
        err = float((grad_k - grad_u).abs().max()) / scale
        # every rank holds the same reduced gradient, bit for bit
        gathered = [torch.empty_like(grad_k) for _ in range(world)]
        dist.all_gather(gathered, grad_k.contiguous())
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        tol = {'fp32': 2e-5, 'bf16x3': 5e-5, 'bf16': 2e-3}[math_mode]
        assert err <= tol, (tag, err)
        assert abs(loss_k - loss_u) <= 1e-5 * max(1.0, abs(loss_u)), (tag, loss_k, loss_u)
        report[tag] = {'grad_max_rel_diff': err, 'loss_k_rank': loss_k, 'loss_1_rank': loss_u}

    # ---- one full learn step: replicas stay identical, and equal the single-GPU step ------------------------------------
    learner.learn_step(mine)
    solo.learn_step(union)
    pb, pt = agent.gn_behaviour.flat_params(), agent.gn_target.flat_params()
    gathered = [torch.empty_like(pb) for _ in range(world)]
    dist.all_gather(gathered, pb.contiguous())
    assert all(torch.equal(g, gathered[0]) for g in gathered), 'replicas diverged after one step'
    gathered = [torch.empty_like(pt) for _ in range(world)]
    dist.all_gather(gathered, pt.contiguous())
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    # the first Adam step moves every weight by lr * g / (|g| + eps): sensitive where |g| ~ eps, so compare where it is not
    big = grad_u.abs() > 1e-6
    dp = float((pb - solo_agent.gn_behaviour.flat_params())[big].abs().max())
    assert dp <= 2e-5, dp
    report['post_adam_max_abs_diff'] = dp
    report.update(world=world, math_mode=math_mode, graphs=ns_all, shard=[lo, hi])
    if rank == 0:
        with open(os.path.join(out_dir, 'mgpu_rank0.json'), 'w') as f:
            json.dump(report, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
