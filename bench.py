#!/usr/bin/env python
"""bench.py -- Q-network fwd+bwd graphs/sec on the replay-minibatch workload of BASELINE.json
(configs[4]: 4096 graphs x 200 vertices, fwd+bwd, gradient all-reduce at N > 1), plus the
standalone PNA aggregation roofline, greedy-rollout graphs/sec (configs[2]) and the CPU
baseline, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one 4096-graph minibatch of synthetic transitions:
graph batching from adjacency bytes, behaviour-net forward with activation stash, TD loss,
full backward into the flat gradient (micro-batched to fit HBM), all-reduce when N > 1.
`value` times it with the minibatch resident in HBM; `e2e` times the same step through the
public API starting from pinned HOST buffers (H2D of adjacency/colours/actions inside the timed
region, loss read back).  Weak scaling: every rank processes its own 4096-graph minibatch.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'qnet_fwd_bwd_graphs_per_sec'
UNIT = 'graphs/s'


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def er_graphs(n_graphs, n, p, seed):
    """Erdos-Renyi G(n,p) adjacency matrices: upper-triangular Bernoulli, mirrored."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n_graphs, n, n), dtype=np.uint8)
    iu = np.triu_indices(n, 1)
    for g in range(n_graphs):
        bits = (rng.random(iu[0].size) < p).astype(np.uint8)
        out[g][iu] = bits
        out[g] = out[g] | out[g].T
    return out


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        # "under load": samples in the upper half of the observed power range
        if sm:
            thr = (min(power) + max(power)) / 2 if power else 0
            load = [s for s, p in zip(sm, power) if p >= thr] or sm
            return {'sm_mhz': float(np.median(load)), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons),
                    'samples': len(sm), 'power_w_max': max(power) if power else None}
        return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}


# ------------------------------------------------------------------------------------------------
def oracle_fwd_bwd_graphs_per_sec(n, sample_graphs, iters, seed=1):
    """The CPU restatement of the reference path (oracle/gn_oracle.py, reference op order, torch CPU
    kernels, all host threads): forward + backward of the DQN loss on `sample_graphs` graphs."""
    import torch
    from oracle import gn_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    adjs = er_graphs(sample_graphs, n, 0.5, seed)
    rng = np.random.default_rng(seed)
    obs = []
    for a in adjs:
        col = -np.ones(n, dtype=np.int64)
        col[rng.permutation(n)[: n // 2]] = rng.integers(0, 8, n // 2)
        obs.append(gn_oracle.observation(a, col))
    x, ei, ea = gn_oracle.collate(obs)
    params = gn_oracle.init_params(seed=0)
    sel = torch.arange(sample_graphs) * n
    tgt = torch.zeros(sample_graphs, 1)
    times = []
    for it in range(iters + 1):
        pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        t0 = time.perf_counter()
        _, q = gn_oracle.gn_forward(pr, x, ei, ea)
        loss = torch.nn.functional.mse_loss(q.view(-1)[sel].view(-1, 1), tgt)
        loss.backward()
        dt = time.perf_counter() - t0
        if it > 0:
            times.append(dt)
    return sample_graphs / float(np.mean(times)), cores, times


_REAL_STDOUT = None


def capture_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first
    communicator, the agent prints 'Starting learning' like the reference): send fd 1 to stderr for the whole run and keep
    the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (the oracle port; the
    reference itself needs torch_geometric/torch_scatter/gym, absent on the box) on host cores."""
    if rank != 0:
        return
    sample = args.cpu_sample_graphs
    t0 = time.perf_counter()
    v, cores, times = oracle_fwd_bwd_graphs_per_sec(args.vertices, sample, max(1, args.steps), seed=1)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': 1, 'ms_per_step': 1e3 * float(np.mean(times)), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'replay minibatch %d graphs x %d vertices fwd+bwd (BASELINE.json configs[4])'
                               % (args.graphs, args.vertices),
                   'sample': '%d graphs x %d vertices per step (bounded sample of the same workload)' % (sample, args.vertices)},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d ER(0.5) graphs x %d vertices, oracle/gn_oracle.py fwd+bwd, torch CPU' % (sample, args.vertices)},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'wall_s': time.perf_counter() - t0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--graphs', type=int, default=4096)
    ap.add_argument('--vertices', type=int, default=200)
    ap.add_argument('--micro-graphs', type=int, default=512)
    ap.add_argument('--math', default='bf16x3', choices=['fp32', 'bf16', 'bf16x3'],
                    help='fp32: CUDA-core FFMA parity path; bf16x3: tcgen05 with hi/lo split operands (near-fp32); bf16: tcgen05')
    ap.add_argument('--rollout-graphs', type=int, default=10000)
    ap.add_argument('--rollout-vertices', type=int, default=50)
    ap.add_argument('--cpu-sample-graphs', type=int, default=8)
    ap.add_argument('--skip-extras', action='store_true', help='only the headline metric')
    args = ap.parse_args()

    capture_stdout()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    os.environ['GCRL_DEVICE'] = 'cuda:%d' % local_rank
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    assert world == args.gpus, 'launch with torchrun --nproc-per-node %d for --gpus %d' % (args.gpus, args.gpus)

    from graph_colouring_with_rl_b200 import _lib, ops
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    from graph_colouring_with_rl_b200.trainer import QLearner, TransitionBatch, synthetic_transitions
    lib = _lib.load()
    hbm_peak, peak_src = peaks()

    # ---- workload: every rank its own minibatch (weak scaling) ------------------------------------
    G, n = args.graphs, args.vertices
    t_gen = time.perf_counter()
    adjs = er_graphs(G, n, 0.5, seed=1 + rank)
    ns, host = synthetic_transitions(list(adjs), 0.5, seed=1 + rank, device=dev)
    assert int(host['action'].min()) >= 0, 'a synthetic episode finished early'
    host = {k: v.pin_memory() for k, v in host.items()}
    log('[bench] rank %d: %d graphs x %d vertices generated in %.1fs' % (rank, G, n, time.perf_counter() - t_gen))
    torch.manual_seed(0)
    agent = DQNAgentGN(2, 1, 0, math_mode=args.math)
    learner = QLearner(agent, max_edges_per_micro_batch=args.micro_graphs * n * (n - 1))
    batch = TransitionBatch.from_host(ns, host, dev)
    q_next = learner.target_q(batch)             # resident TD targets (the target forward is not part of fwd+bwd)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / steps

    log('[bench] rank %d: agent + batch ready' % rank)
    # ---- headline: resident inputs ------------------------------------------------------------------
    step_resident = lambda: learner.fwd_bwd(batch, q_next)
    for _ in range(args.warmup):
        step_resident()
    lib.gcrl_prof_enable(1)
    import ctypes
    ms_k = (ctypes.c_double * 8)()
    cnt_k = (ctypes.c_int64 * 8)()
    lib.gcrl_prof_read(ms_k, cnt_k, 8)            # drop warm-up records
    launches0 = lib.gcrl_launch_count()
    clocks = ClockSampler(local_rank)
    torch.cuda.profiler.start()                   # `ncu --profile-from-start off` sees exactly the timed steps
    ms_step = timed(step_resident, args.steps)
    torch.cuda.profiler.stop()
    launches = lib.gcrl_launch_count() - launches0
    lib.gcrl_prof_read(ms_k, cnt_k, 8)
    lib.gcrl_prof_enable(0)
    value = G * world / (ms_step * 1e-3)
    log('[bench] rank %d: resident %.1f ms/step' % (rank, ms_step))

    # ---- e2e: host buffers -> public API -> loss on the host -------------------------------------------
    def step_e2e():
        b = TransitionBatch.from_host(ns, host, dev)
        loss = learner.fwd_bwd(b, q_next)
        return float(loss.item())
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clk = clocks.stop()
    e2e_value = G * world / (ms_e2e * 1e-3)
    log('[bench] rank %d: e2e %.1f ms/step' % (rank, ms_e2e))
    h2d = TransitionBatch.host_bytes(host)

    # ---- roofline of the dominant kernel (edge backward, blocks 2-4) -------------------------------------
    kinds = ['edge_fwd_first', 'edge_fwd', 'edge_fwd_last', 'edge_bwd_first', 'edge_bwd', 'edge_bwd_last', 'pna_fwd', 'pna_bwd']
    # block 5: the tensor-core modes stash e_5 (forward writes it: counted under edge_fwd) and its backward then reads
    # e_4, e_5 and writes d e_4 (768 B/edge); the recompute flavour (fp32, GCRL_EDGE_IMPL=tc) reads e_4, writes d e_4
    recompute_e5 = args.math == 'fp32' or os.environ.get('GCRL_EDGE_IMPL', '') == 'tc'
    bytes_per_edge = {'edge_fwd_first': 4 + 256, 'edge_fwd': 512, 'edge_fwd_last': 256, 'edge_bwd_first': 4 + 512,
                      'edge_bwd': 1024, 'edge_bwd_last': 512 if recompute_e5 else 768}
    E_micro = min(args.micro_graphs, G) * n * (n - 1)
    kernels = {}
    for i, k in enumerate(kinds):
        if cnt_k[i]:
            avg = ms_k[i] / cnt_k[i]
            kernels[k] = {'launches_per_step': cnt_k[i] / args.steps, 'avg_ms': avg, 'share_of_step': ms_k[i] / args.steps / ms_step}
            if k in bytes_per_edge:
                kernels[k]['gbps'] = bytes_per_edge[k] * E_micro / (avg * 1e-3) / 1e9
    dom = max((k for k in kernels if k in bytes_per_edge), key=lambda k: kernels[k]['share_of_step'])
    # which kernel serves each launch kind (gn_forward.cu / gn_backward.cu dispatch): the multi-group kernels
    # (k_edge_fwd_wg / k_edge_bwd_wg) unless GCRL_FWD_IMPL / GCRL_EDGE_IMPL select an older one
    fwd_impl = {'ws': '_ws', 'tc': '_tc'}.get(os.environ.get('GCRL_FWD_IMPL', '')[:2], '_wg')
    bwd_impl = {'ws': '_ws', 'tc': '_tc'}.get(os.environ.get('GCRL_EDGE_IMPL', '')[:2], '_wg')
    if args.math == 'fp32':
        fwd_impl = bwd_impl = ''
    kname = {'edge_fwd_first': 'k_edge_fwd%s<first>' % fwd_impl, 'edge_fwd': 'k_edge_fwd%s' % fwd_impl,
             'edge_fwd_last': 'k_edge_fwd%s<last>' % fwd_impl, 'edge_bwd_first': 'k_edge_bwd%s<first>' % bwd_impl,
             'edge_bwd': 'k_edge_bwd%s' % bwd_impl,
             'edge_bwd_last': 'k_edge_bwd%s<block 5>' % bwd_impl}[dom]
    # DRAM traffic per launch: bytes per edge of this kernel from the committed ncu capture
    # (profiles/r01b_traffic.json, dram__bytes_read.sum + dram__bytes_write.sum over the edges of that launch)
    traffic = None
    try:
        tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles',
                                         'r01b_traffic.json' if bwd_impl == '_wg' else 'r01_traffic.json')))
        ent = tj.get(args.math, {}).get(dom)
        if ent:
            traffic = ent['dram_bytes_per_edge'] * E_micro
    except (OSError, ValueError):
        pass
    roofline = {'bound': 'hbm', 'kernel': kname, 'achieved': kernels[dom]['gbps'], 'peak': hbm_peak,
                'unit': 'GB/s', 'frac': kernels[dom]['gbps'] / hbm_peak, 'traffic': traffic, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': bytes_per_edge[dom] * E_micro,
                'share_of_step': kernels[dom]['share_of_step']}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': {'fp32': 'f32', 'bf16x3': 'bf16x3 (hi+lo split operands, f32 accumulate)', 'bf16': 'bf16 (f32 accumulate)'}[args.math], 'data': 'synthetic',
        'config': {'workload': 'replay minibatch %d graphs x %d vertices fwd+bwd per GPU, ER(0.5), states 50%% coloured '
                               '(BASELINE.json configs[4])' % (G, n),
                   'global_graphs': G * world, 'edges_per_gpu': int(batch.n_edges), 'micro_batch_graphs': args.micro_graphs,
                   'math': args.math, 'parallelism': 'dp%d (graphs sharded, 794 KB grad all-reduce)' % world,
                   'l2': 'inputs larger than L2 (%.1f GB edge state per micro-batch tensor)' % (E_micro * 256 / 1e9)},
        'roofline': roofline, 'kernels': kernels, 'gpu_launches': int(launches),
        'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4},
        'clocks': clk,
    }

    if not args.skip_extras:
        line['learn_step'] = bench_learn_step(learner, batch, timed, args, G, world)   # all ranks (all-reduce inside)
    if not args.skip_extras:
        # free the big buffers before the side measurements
        del batch
        learner._stash = None
        _lib._ws_cache.clear()
        torch.cuda.empty_cache()
        # single-GPU side measurements and the CPU baseline: at N = 1 only (at N > 1 the other ranks would idle at the
        # final barrier while rank 0 runs them)
        single = rank == 0 and world == 1
        if single:
            line['pna'] = bench_pna(lib, ops, dev, hbm_peak, peak_src, args)
        line_roll = bench_rollouts(agent, dev, args, rank, world, max_over_ranks, barrier)
        if rank == 0:
            line['rollouts'] = line_roll
        if single:
            line['train_loop'] = bench_train_loop(dev, args)
            line['large_graphs'] = bench_large_graphs(agent, dev)
            v, cores, times = oracle_fwd_bwd_graphs_per_sec(n, args.cpu_sample_graphs, 3)
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                    'sample': '%d ER(0.5) graphs x %d vertices, oracle/gn_oracle.py fwd+bwd (reference op order, '
                                              'torch CPU, %d threads), mean of 3 after 1 warm-up' % (args.cpu_sample_graphs, n, cores)}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_train_loop(dev, args):
    """The device training loop (dqn_main.DeviceTrainer, reference dqn_main.py:158-199) on the shapes of
    BASELINE.json configs[1]: graphs of 15-50 vertices, minibatch 64, learn every 5 environment steps, 64 episodes
    in lock step, everything per-step on the device (replay ring included)."""
    import contextlib
    import numpy as np
    import torch
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    from graph_colouring_with_rl_b200.dqn_main import DeviceTrainer
    from graph_colouring_with_rl_b200.utils import erdos_renyi_adjacency
    rng = np.random.default_rng(0)
    graphs = [erdos_renyi_adjacency(int(n), 0.5, rng) for n in rng.integers(15, 51, size=256)]
    torch.manual_seed(0)
    agent = DQNAgentGN(2, 1, 0, math_mode=args.math)
    n_envs, warm, timed_eps = 64, 128, 512
    with contextlib.redirect_stdout(sys.stderr):          # the agent prints 'Starting learning' like the reference
        tr = DeviceTrainer(agent, graphs, None, n_envs=n_envs, device_seed=0)
        tr.run(warm, verbose=False)                       # fills the ring past the 10 x batch gate
        torch.cuda.synchronize()
        t0, s0, l0 = time.perf_counter(), tr.t_step, len(tr.losses)
        tr.run(warm + timed_eps, verbose=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return {'workload': '256 ER(0.5) graphs of 15-50 vertices, %d episodes in lock step, minibatch %d, learn every %d env steps, '
                        '%d timed episodes after %d warm-up' % (n_envs, agent.batch_size, agent.update_every, timed_eps, warm),
            'episodes_per_sec': timed_eps / dt, 'env_steps_per_sec': (tr.t_step - s0) / dt,
            'learn_steps_per_sec': (len(tr.losses) - l0) / dt, 'replay_ring_bytes': tr.memory.bytes(),
            'timing': 'host wall clock around the loop (it synchronises once per environment step)'}


def bench_learn_step(learner, batch, timed, args, G, world):
    """Full DQN learn step: target forward + fwd/bwd + fused Adam/soft update."""
    learner.learn_step(batch)
    ms = timed(lambda: learner.learn_step(batch), max(1, args.steps // 2))
    return {'steps_per_sec': 1e3 / ms, 'transitions_per_sec': G * world * 1e3 / ms, 'ms_per_step': ms}


def bench_pna(lib, ops, dev, hbm_peak, peak_src, args):
    """Standalone PNA aggregation on one block of the rollout config (10k graphs x 50 vertices):
    [E,64] fp32 messages -> [N,256]; algorithmic bytes 4*64*E + 4*256*N + 4*(N+1)."""
    import ctypes
    import torch
    Gr, nr = args.rollout_graphs, args.rollout_vertices
    N, E = Gr * nr, Gr * nr * (nr - 1)
    seg = (torch.arange(N + 1, device=dev, dtype=torch.int64) * (nr - 1)).int()
    msg = torch.randn(E, 64, device=dev)
    for _ in range(3):
        ops.pna_forward(msg, seg, N)
    torch.cuda.synchronize()
    ms_k = (ctypes.c_double * 8)()
    cnt_k = (ctypes.c_int64 * 8)()
    lib.gcrl_prof_enable(1)
    lib.gcrl_prof_read(ms_k, cnt_k, 8)
    for _ in range(10):
        ops.pna_forward(msg, seg, N)
    lib.gcrl_prof_read(ms_k, cnt_k, 8)
    lib.gcrl_prof_enable(0)
    avg = ms_k[6] / max(cnt_k[6], 1)
    nbytes = 4 * 64 * E + 4 * 256 * N + 4 * (N + 1)
    gbps = nbytes / (avg * 1e-3) / 1e9
    return {'workload': '%d graphs x %d vertices, one block: msg [%d,64] fp32' % (Gr, nr, E), 'avg_ms': avg,
            'algorithmic_bytes': nbytes, 'achieved': gbps, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbps / hbm_peak,
            'peak_source': peak_src, 'l2': 'input 6.3 GB >> L2'}


def bench_large_graphs(agent, dev):
    """Larger-graph inference (BASELINE.json configs[3]): one action-scoring step (Q of every vertex + masked argmax) on a
    single ER(0.5) graph of 100-1000 vertices (up to 999 000 complete-graph edges), and on a batch of 32 x 250 vertices."""
    import torch
    from graph_colouring_with_rl_b200 import ops
    from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
    from graph_colouring_with_rl_b200.utils import erdos_renyi_adjacency
    net = agent.gn_behaviour
    flat = net.flat_params()
    rng = np.random.default_rng(3)
    out = {}
    for tag, n, B in (('n100', 100, 1), ('n250', 250, 1), ('n500', 500, 1), ('n1000', 1000, 1), ('32xn250', 250, 32)):
        env = BatchedColouringEnv([erdos_renyi_adjacency(n, 0.5, rng) for _ in range(B)], device=dev)
        packed, eattr = env.packed
        env.step(env.first_vertices())

        def step():
            _, q, _ = ops.gn_forward(flat, net.dims, packed, env.x, eattr, math_mode=net.math_mode)
            return ops.score_argmax(q, env.colour, env.node_off)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            step()
        b.record()
        torch.cuda.synchronize()
        out[tag] = {'ms_per_scoring_step': a.elapsed_time(b) / 20, 'edges': int(env.n_edges)}
    out['workload'] = 'Q of every vertex + masked argmax per call, ER(0.5), mean of 20 calls after 3 warm-ups'
    return out


def bench_rollouts(agent, dev, args, rank, world, max_over_ranks, barrier):
    """Greedy colouring rollouts (BASELINE.json configs[2]): rollout_graphs ER(0.5) graphs of
    rollout_vertices vertices, contiguous shards per rank, full episodes on the device."""
    import torch
    from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
    from graph_colouring_with_rl_b200.rollouts import greedy_rollouts
    Gr, nr = args.rollout_graphs, args.rollout_vertices
    lo, hi = Gr * rank // world, Gr * (rank + 1) // world
    adjs = er_graphs(Gr, nr, 0.5, seed=0)[lo:hi]
    env = BatchedColouringEnv(list(adjs), device=dev)
    greedy_rollouts(None, agent.gn_behaviour, env=env)             # warm-up sweep
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    res = greedy_rollouts(None, agent.gn_behaviour, env=env)
    b.record()
    barrier()
    ms = max_over_ranks(a.elapsed_time(b))
    used = res['colours_used'].float().mean().item()
    return {'graphs_per_sec': Gr / (ms * 1e-3), 'ms_per_sweep': ms, 'steps': res['steps'],
            'workload': '%d ER(0.5) graphs x %d vertices, greedy episodes, default-init weights' % (Gr, nr),
            'mean_colours_used_rank0': used}


if __name__ == '__main__':
    main()
