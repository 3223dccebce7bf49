#!/usr/bin/env python
"""bench.py -- Q-network fwd+bwd graphs/sec on the replay-minibatch workload of BASELINE.json
(configs[4]: ONE minibatch of 4096 graphs x 200 vertices, fwd+bwd, split by graph over the N GPUs with one
gradient all-reduce), plus greedy-rollout graphs/sec (configs[2]), the standalone PNA aggregation roofline, the
reference-shaped DQN learn step (configs[1]) and the CPU baselines; one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over the 4096-graph minibatch of synthetic transitions: graph batching from
adjacency bytes, behaviour-net forward with activation stash, TD loss, full backward into the flat gradient
(micro-batched to fit HBM), ONE asynchronous all-reduce of [gradient | loss | graph count] when N > 1.
STRONG scaling: the same global minibatch at every N (rank r holds graphs [G r / N, G (r+1) / N)), so `value` at
N GPUs is G / (time of the slowest rank) and the loss / gradient checksum in `check` is the same number at every N.
`value` times the step with the rank's shard resident in HBM; `e2e` times it through the public API starting from
pinned HOST buffers (H2D of adjacency / colours / actions inside the timed region, loss read back).
"""
import argparse
import contextlib
import io
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
CHUNK = 512          # graphs per generation chunk: the global minibatch is the same at every N


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


# ---- CPU baselines: the reference's own modules (oracle/_ref, kind "reference") or the oracle port ---------------------
def _load_reference():
    """The unmodified reference (oracle/ref_import.py: /root/reference here, its byte-for-byte copy oracle/_ref on the
    GPU box, third-party stand-ins from oracle/shim) or None.  Must run before torch is imported with a GPU visible:
    architecture.py:11 pins cuda:4, so the reference is only ever run with CUDA hidden from its process."""
    try:
        from oracle import ref_import
        if not ref_import.reference_available():
            return None, None
        with contextlib.redirect_stdout(sys.stderr):
            return ref_import.import_reference(), ref_import
    except Exception as e:          # noqa: BLE001 -- any import problem falls back to the port, and says so
        log('[bench] reference import failed (%r): falling back to the oracle port' % (e,))
        return None, None


def cpu_fwd_bwd_graphs_per_sec(n, sample_graphs, iters, warmup=1, seed=1, ref=None):
    """Forward + backward of the DQN loss on `sample_graphs` ER(0.5) graphs of n vertices on the host cores.
    With the reference loaded: its own GNNetwork.forward (architecture.py:179-199) over Batch.from_data_list, F.mse_loss
    on the action vertices, optimizer.zero_grad() + loss.backward() (dqn_agent.py:174-219).  Otherwise the oracle port
    (oracle/gn_oracle.py, reference op order on torch CPU kernels)."""
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
    sel = torch.arange(sample_graphs) * n
    tgt = torch.zeros(sample_graphs, 1)
    times = []
    if ref is not None:
        from torch_geometric.data import Batch, Data       # oracle/shim stand-ins
        torch.manual_seed(0)
        net = ref.architecture.GNNetwork(2, 1, 0, 'GN', 1e-3)
        batch = Batch.from_data_list([Data(x=x, edge_index=ei, edge_attr=ea) for x, ei, ea in obs])
        vmap = [g for g in range(sample_graphs) for _ in range(n)]
        gf = torch.zeros(sample_graphs, 0)
        for it in range(iters + warmup):
            t0 = time.perf_counter()
            _, q = net(batch, gf, vmap, None)
            loss = torch.nn.functional.mse_loss(q.view(-1)[sel].view(-1, 1), tgt)
            net.optimizer.zero_grad()
            loss.backward()
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
        kind = 'reference'
    else:
        x, ei, ea = gn_oracle.collate(obs)
        params = gn_oracle.init_params(seed=0)
        for it in range(iters + warmup):
            pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
            t0 = time.perf_counter()
            _, q = gn_oracle.gn_forward(pr, x, ei, ea)
            loss = torch.nn.functional.mse_loss(q.view(-1)[sel].view(-1, 1), tgt)
            loss.backward()
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
        kind = 'port'
    return sample_graphs / float(np.mean(times)), cores, times, kind


def cpu_rollouts_graphs_per_sec(n, sample_graphs, ref=None, ref_import=None, seed=0):
    """Greedy colouring episodes of `sample_graphs` ER(0.5) graphs on the host: the reference's own
    test_using_learned_policy (test_policy_functions.py:6-103: its environment, its agent.choose_action) or the port
    (oracle/env_oracle.greedy_rollout over oracle/gn_oracle.gn_forward)."""
    import torch
    from oracle import env_oracle, gn_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    adjs = er_graphs(sample_graphs, n, 0.5, seed)
    if ref is not None:
        targets = [ref_import.reference_target(ref, a, 'g%d' % i) for i, a in enumerate(adjs)]
        torch.manual_seed(0)
        agent = ref.dqn_agent.DQNAgentGN(2, 1, 0, test_mode=True)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            ref.test_policy_functions.test_using_learned_policy(targets, agent, False)
        dt = time.perf_counter() - t0
        kind = 'reference'
    else:
        params = gn_oracle.init_params(seed=0)

        def q_of(a):
            def f(col):
                x, ei, ea = gn_oracle.observation(a, col)
                with torch.no_grad():
                    return gn_oracle.gn_forward(params, x, ei, ea)[1].view(-1).numpy()
            return f
        t0 = time.perf_counter()
        for a in adjs:
            env_oracle.greedy_rollout(a, q_of(a))
        dt = time.perf_counter() - t0
        kind = 'port'
    return sample_graphs / dt, cores, dt, kind


def cpu_learn_steps_per_sec(ref, ref_import, iters=3, seed=0):
    """The reference's DQNAgentGN.learn() (dqn_agent.py:147-223: sample 64 of the stored transitions, two forwards, TD
    target, MSE, backward, Adam, soft update) on transitions of ER(0.5) graphs of 15-50 vertices gathered by random
    play in the reference's own environment (BASELINE.json configs[1] shapes).  Needs the reference itself."""
    import random
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    adjs = [er_graphs(1, int(n), 0.5, 100 + i)[0] for i, n in enumerate(rng.integers(15, 51, size=24))]
    targets = [ref_import.reference_target(ref, a, 'g%d' % i) for i, a in enumerate(adjs)]
    with contextlib.redirect_stdout(io.StringIO()):
        env = ref.env.GraphColouring(dataset_graphs=targets)
        agent = ref.dqn_agent.DQNAgentGN(2, 1, 0)
        need = 10 * agent.batch_size
        while agent.memory.current_mem_size < need:
            _, target, cur = env.reset()
            data, gf = env.GenerateData_v1(target, cur)
            done = False
            while not done and agent.memory.current_mem_size < need:
                v = random.choice(cur['unassigned_vertices'])
                nxt, reward, done, _ = env.step(v)
                ndata, ngf = env.GenerateData_v1(target, nxt)
                agent.remember(data, gf, nxt['assigned_vertices_onehot'], v, reward, ndata, ngf, done)
                data, gf, cur = ndata, ngf, nxt
        agent.learn()                                           # warm-up
        t0 = time.perf_counter()
        for _ in range(iters):
            agent.learn()
        dt = (time.perf_counter() - t0) / iters
    return 1.0 / dt, cores, dt


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


def workload_name(G, n):
    return ('replay minibatch %d graphs x %d vertices fwd+bwd, ER(0.5), states 50%% coloured (BASELINE.json configs[4])'
            % (G, n))


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (its unmodified modules from oracle/_ref
    under the third-party stand-ins of oracle/shim; the oracle port if that copy is absent) on the host cores, on a
    bounded sample of the same workload per step.  Rank 0 alone runs it."""
    if rank != 0:
        return
    os.environ['CUDA_VISIBLE_DEVICES'] = ''                # architecture.py:11 pins cuda:4: the reference runs on the CPU
    ref, _ = _load_reference()
    sample = args.cpu_sample_graphs
    t0 = time.perf_counter()
    v, cores, times, kind = cpu_fwd_bwd_graphs_per_sec(args.vertices, sample, max(1, args.steps), warmup=max(1, args.warmup),
                                                       seed=1, ref=ref)
    what = ("the reference's own GNNetwork.forward + mse_loss + backward (oracle/_ref, unmodified files; torch_geometric / "
            "torch_scatter stand-ins from oracle/shim)" if kind == 'reference' else 'oracle/gn_oracle.py fwd+bwd (port)')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': max(1, args.warmup), 'ms_per_step': 1e3 * float(np.mean(times)), 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.graphs, args.vertices),
                   'sample': '%d graphs x %d vertices per step (bounded sample of the same workload)' % (sample, args.vertices)},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind,
                         'sample': '%d ER(0.5) graphs x %d vertices per step, %s, torch CPU, %d threads' % (sample, args.vertices, what, cores)},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'wall_s': time.perf_counter() - t0,
    }
    emit(line)


def global_shard(G, n, lo, hi, dev, seed_base=1):
    """Graphs [lo, hi) of THE global minibatch: the minibatch is defined chunk by chunk (CHUNK graphs: adjacency seed
    1000 + c, transition seed seed_base + c), so every N sees the same 4096 transitions whatever its split."""
    from graph_colouring_with_rl_b200.trainer import FIELDS, shard_host, synthetic_transitions
    import torch
    parts_ns, parts = [], []
    c0, c1 = lo // CHUNK, (max(hi, lo + 1) - 1) // CHUNK + 1
    for c in range(c0, c1):
        g0, g1 = c * CHUNK, min(G, (c + 1) * CHUNK)
        adjs = er_graphs(g1 - g0, n, 0.5, seed=1000 + c)
        ns, host = synthetic_transitions(list(adjs), 0.5, seed=seed_base + c, device=dev)
        a, b = max(lo, g0) - g0, min(hi, g1) - g0
        if a < b:
            ns, host = shard_host(ns, host, a, b)
            parts_ns.append(ns)
            parts.append(host)
    ns = np.concatenate(parts_ns) if parts_ns else np.zeros(0, dtype=np.int64)
    host = {k: torch.cat([p[k] for p in parts]) for k in FIELDS}
    return ns, host


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--graphs', type=int, default=4096, help='graphs of the GLOBAL minibatch (split over the GPUs)')
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
    from graph_colouring_with_rl_b200.trainer import QLearner, TransitionBatch, shard_range
    lib = _lib.load()
    hbm_peak, peak_src = peaks()

    # ---- workload: ONE global minibatch, this rank's contiguous shard of it (strong scaling) ------------------------------
    G, n = args.graphs, args.vertices
    lo, hi = shard_range(G, rank, world)
    t_gen = time.perf_counter()
    ns, host = global_shard(G, n, lo, hi, dev)
    assert int(host['action'].min()) >= 0, 'a synthetic episode finished early'
    host = {k: v.pin_memory() for k, v in host.items()}
    log('[bench] rank %d: graphs [%d, %d) of %d x %d vertices generated in %.1fs' % (rank, lo, hi, G, n, time.perf_counter() - t_gen))
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
    step_resident = lambda: learner.fwd_bwd(batch, q_next, batch_total=G)
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
    value = G / (ms_step * 1e-3)
    # what the timed steps computed: the global loss and a checksum of the (all-reduced) gradient -- the same numbers at
    # every N because the global minibatch is the same (tests/test_headline_gpu.py checks them against the oracle)
    loss_v = float(learner._gbuf[-2].item())
    grad = agent.gn_behaviour._flat_grad
    check = {'loss': loss_v, 'grad_l2': float(grad.double().norm().item()), 'grad_sum': float(grad.double().sum().item()),
             'grad_abs_max': float(grad.abs().max().item())}
    assert all(np.isfinite(v) for v in check.values()) and check['grad_l2'] > 0, check
    log('[bench] rank %d: resident %.1f ms/step, loss %.6g, |grad| %.6g' % (rank, ms_step, check['loss'], check['grad_l2']))

    # ---- e2e: host buffers -> public API -> loss on the host -------------------------------------------
    def step_e2e():
        b = TransitionBatch.from_host(ns, host, dev)
        loss = learner.fwd_bwd(b, q_next, batch_total=G)
        return float(loss.item())
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clk = clocks.stop()
    e2e_value = G / (ms_e2e * 1e-3)
    log('[bench] rank %d: e2e %.1f ms/step' % (rank, ms_e2e))
    h2d = TransitionBatch.host_bytes(host)

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------------------
    kinds = ['edge_fwd_first', 'edge_fwd', 'edge_fwd_last', 'edge_bwd_first', 'edge_bwd', 'edge_bwd_last', 'pna_fwd', 'pna_bwd']
    # block 5: the tensor-core modes stash e_5 (forward writes it: counted under edge_fwd) and its backward then reads
    # e_4, e_5 and writes d e_4 (768 B/edge); the fp32 flavour recomputes e_5 (reads e_4, writes d e_4)
    recompute_e5 = args.math == 'fp32'
    bytes_per_edge = {'edge_fwd_first': 4 + 256, 'edge_fwd': 512, 'edge_fwd_last': 256, 'edge_bwd_first': 4 + 512,
                      'edge_bwd': 1024, 'edge_bwd_last': 512 if recompute_e5 else 768}
    E_micro = min(args.micro_graphs, hi - lo) * n * (n - 1)
    kernels = {}
    for i, k in enumerate(kinds):
        if cnt_k[i]:
            avg = ms_k[i] / cnt_k[i]
            kernels[k] = {'launches_per_step': cnt_k[i] / args.steps, 'avg_ms': avg, 'share_of_step': ms_k[i] / args.steps / ms_step}
            if k in bytes_per_edge:
                kernels[k]['gbps'] = bytes_per_edge[k] * E_micro / (avg * 1e-3) / 1e9
                kernels[k]['frac_of_hbm_peak'] = kernels[k]['gbps'] / hbm_peak
    dom = max((k for k in kernels if k in bytes_per_edge), key=lambda k: kernels[k]['share_of_step'])
    sfx = '' if args.math == 'fp32' else '_wg'
    kname = {'edge_fwd_first': 'k_edge_fwd%s<first>' % sfx, 'edge_fwd': 'k_edge_fwd%s' % sfx,
             'edge_fwd_last': 'k_edge_fwd%s<last>' % sfx, 'edge_bwd_first': 'k_edge_bwd%s<first>' % sfx,
             'edge_bwd': 'k_edge_bwd%s' % sfx, 'edge_bwd_last': 'k_edge_bwd%s<block 5>' % sfx}[dom]
    # DRAM traffic per launch: bytes per edge of this kernel from the committed ncu capture
    # (profiles/*_traffic.json, dram__bytes_read.sum + dram__bytes_write.sum over the edges of that launch)
    traffic = None
    for name in ('r02_traffic.json', 'r01b_traffic.json'):
        try:
            tj = json.load(open(os.path.join(ROOT, 'profiles', name)))
            ent = tj.get(args.math, {}).get(dom)
            if ent:
                traffic = ent['dram_bytes_per_edge'] * E_micro
                break
        except (OSError, ValueError):
            pass
    edge_bytes_step = sum(bytes_per_edge[k] * E_micro * kernels[k]['launches_per_step'] for k in kernels if k in bytes_per_edge)
    roofline = {'bound': 'hbm', 'kernel': kname, 'achieved': kernels[dom]['gbps'], 'peak': hbm_peak,
                'unit': 'GB/s', 'frac': kernels[dom]['gbps'] / hbm_peak, 'traffic': traffic, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': bytes_per_edge[dom] * E_micro,
                'share_of_step': kernels[dom]['share_of_step'],
                'whole_step': {'algorithmic_edge_bytes': edge_bytes_step, 'gbps': edge_bytes_step / (ms_step * 1e-3) / 1e9,
                               'frac': edge_bytes_step / (ms_step * 1e-3) / 1e9 / hbm_peak}}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': {'fp32': 'f32', 'bf16x3': 'bf16x3 (hi+lo split operands, f32 accumulate)', 'bf16': 'bf16 (f32 accumulate)'}[args.math], 'data': 'synthetic',
        'config': {'workload': workload_name(G, n),
                   'global_graphs': G, 'graphs_per_gpu': hi - lo, 'edges_per_gpu': int(batch.n_edges),
                   'micro_batch_graphs': min(args.micro_graphs, hi - lo), 'math': args.math,
                   'parallelism': 'dp%d: ONE minibatch split by graph, one async all-reduce of [794 KB gradient | loss | count]' % world,
                   'l2': 'inputs larger than L2 (%.1f GB edge state per micro-batch tensor)' % (E_micro * 256 / 1e9)},
        'learn_steps_per_sec_fwd_bwd_only': 1e3 / ms_step,
        'check': check,
        'roofline': roofline, 'kernels': kernels, 'gpu_launches': int(launches),
        'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': h2d * world if world > 1 else h2d,
                'd2h_bytes_per_step': 4 * world},
        'clocks': clk,
    }

    if not args.skip_extras:
        line['learn_step'] = bench_learn_step(learner, batch, timed, args, G)          # all ranks (all-reduce inside)
    if not args.skip_extras:
        # free the big buffers before the side measurements
        del batch
        learner._stash = None
        _lib._ws_cache.clear()
        torch.cuda.empty_cache()
        if world > 1:
            line['weak_scaling'] = bench_weak(agent, args, rank, world, dev, timed)
        single = rank == 0 and world == 1
        if single:
            line['value_fp32'] = bench_fp32(args, ns, host, dev)
            line['pna'] = bench_pna(lib, ops, dev, hbm_peak, peak_src, args)
        _lib._ws_cache.clear()
        torch.cuda.empty_cache()
        line_roll = bench_rollouts(agent, dev, args, rank, world, max_over_ranks, barrier, lib, hbm_peak)
        if rank == 0:
            line['rollouts'] = line_roll
        if single:
            # single-GPU side measurements and the CPU baselines: at N = 1 only (at N > 1 the other ranks would idle at the
            # final barrier while rank 0 runs them)
            line['train_loop'] = bench_train_loop(dev, args)
            line['dropin'] = bench_dropin(dev, args)
            line['large_graphs'] = bench_large_graphs(agent, dev)
            line.update(bench_cpu_baselines(args, n))
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_cpu_baselines(args, n):
    """The CPU arm of every half of the metric, in a SUBPROCESS with CUDA hidden (the reference pins cuda:4): fwd+bwd
    graphs/s, rollout graphs/s and learn steps/s of the reference's own modules on the host cores."""
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--cpu-legs', '--vertices', str(n),
           '--cpu-sample-graphs', str(args.cpu_sample_graphs), '--rollout-vertices', str(args.rollout_vertices)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900, env=env)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:      # noqa: BLE001
        return {'cpu_baseline': {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port', 'sample': 'failed: %r' % (e,)}}


def cpu_legs(args):
    """Child of bench_cpu_baselines: prints {'cpu_baseline', 'rollouts_cpu_baseline', 'learn_step_cpu_baseline'}."""
    capture_stdout()
    ref, ref_import = _load_reference()
    n = args.vertices
    out = {}
    v, cores, times, kind = cpu_fwd_bwd_graphs_per_sec(n, args.cpu_sample_graphs, 3, warmup=1, ref=ref)
    out['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind,
                           'sample': '%d ER(0.5) graphs x %d vertices, %s fwd+bwd on torch CPU (%d threads), mean of 3 after 1 warm-up'
                                     % (args.cpu_sample_graphs, n, "the reference's own GNNetwork (oracle/_ref)" if kind == 'reference'
                                        else 'oracle/gn_oracle.py (port)', cores)}
    rg = 6
    v, cores, dt, kind = cpu_rollouts_graphs_per_sec(args.rollout_vertices, rg, ref, ref_import)
    out['rollouts_cpu_baseline'] = {'value': v, 'unit': 'graphs/s', 'cores': cores, 'kind': kind,
                                    'sample': '%d ER(0.5) graphs x %d vertices, greedy episodes by %s, %.1f s' % (
                                        rg, args.rollout_vertices, "the reference's test_using_learned_policy (its environment, its "
                                        "agent.choose_action)" if kind == 'reference' else 'oracle/env_oracle.greedy_rollout (port)', dt)}
    if ref is not None:
        v, cores, dt = cpu_learn_steps_per_sec(ref, ref_import)
        out['learn_step_cpu_baseline'] = {'value': v, 'unit': 'learn steps/s', 'cores': cores, 'kind': 'reference',
                                          'sample': "the reference's DQNAgentGN.learn() on 64 of 640 stored transitions of ER(0.5) graphs "
                                                    'of 15-50 vertices, mean of 3 after 1 warm-up (%.3f s/step)' % dt}
    emit(out)


def bench_weak(agent, args, rank, world, dev, timed):
    """Weak scaling extra (round 1's headline): every rank its OWN minibatch of --graphs graphs, global minibatch
    N x larger, the same single all-reduce."""
    import torch
    from graph_colouring_with_rl_b200.trainer import QLearner, TransitionBatch
    G, n = args.graphs, args.vertices
    ns, host = global_shard(G, n, 0, G, dev, seed_base=5000 + 100 * rank)
    batch = TransitionBatch.from_host(ns, host, dev)
    learner = QLearner(agent, max_edges_per_micro_batch=args.micro_graphs * n * (n - 1), sync_params=False)
    q_next = learner.target_q(batch)
    step = lambda: learner.fwd_bwd(batch, q_next, batch_total=G * world)
    step()
    steps = max(2, args.steps // 4)
    ms = timed(step, steps)
    del batch
    learner._stash = None
    torch.cuda.empty_cache()
    return {'value': G * world / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'graphs_per_gpu': G, 'global_graphs': G * world,
            'steps': steps, 'scaling': 'weak'}


def bench_fp32(args, ns, host, dev):
    """The same minibatch in math mode fp32 (CUDA-core FFMA: the bit-level parity path, same precision as the reference)."""
    import torch
    from graph_colouring_with_rl_b200 import _lib
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    from graph_colouring_with_rl_b200.trainer import QLearner, TransitionBatch
    G, n = args.graphs, args.vertices
    torch.manual_seed(0)
    agent = DQNAgentGN(2, 1, 0, math_mode='fp32')
    learner = QLearner(agent, max_edges_per_micro_batch=args.micro_graphs * n * (n - 1), local_only=True)
    batch = TransitionBatch.from_host(ns, host, dev)
    q_next = learner.target_q(batch)
    learner.fwd_bwd(batch, q_next)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    steps = 2
    for _ in range(steps):
        learner.fwd_bwd(batch, q_next)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    loss = float(learner._gbuf[-2].item())
    del batch
    learner._stash = None
    _lib._ws_cache.clear()
    torch.cuda.empty_cache()
    return {'value': G / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'steps': steps, 'warmup': 1, 'dtype': 'f32', 'loss': loss,
            'note': 'math mode fp32 (true fp32 FFMA, the precision of the reference); the headline is bf16x3'}


def bench_train_loop(dev, args):
    """The device training loop (dqn_main.DeviceTrainer, reference dqn_main.py:158-199) on the shapes of
    BASELINE.json configs[1]: graphs of 15-50 vertices, minibatch 64, learn every 5 environment steps, 64 episodes
    in lock step, everything per-step on the device (replay ring included)."""
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
        # one learn step alone, device-timed
        for _ in range(3):
            tr.learn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            tr.learn()
        b.record()
        torch.cuda.synchronize()
        learn_ms = a.elapsed_time(b) / 20
    return {'workload': '256 ER(0.5) graphs of 15-50 vertices, %d episodes in lock step, minibatch %d, learn every %d env steps, '
                        '%d timed episodes after %d warm-up' % (n_envs, agent.batch_size, agent.update_every, timed_eps, warm),
            'episodes_per_sec': timed_eps / dt, 'env_steps_per_sec': (tr.t_step - s0) / dt,
            'learn_steps_per_sec': (len(tr.losses) - l0) / dt, 'learn_step_ms': learn_ms, 'replay_ring_bytes': tr.memory.bytes(),
            'note': 'a 64-transition learn step is ~70 k edges: latency-bound (launch chain), not bandwidth-bound',
            'timing': 'host wall clock around the loop (it synchronises once per environment step); learn_step_ms: CUDA events'}


def bench_dropin(dev, args):
    """The reference-signature path, timed once: `agent.remember(Data...)` / `agent.learn()` on 64
    transitions of 15-50-vertex graphs -- with the environment's own observations (dense-adjacency batching) and with
    plain Data objects carrying an int64 edge_index (16 E bytes through gcrl_pack) -- and batch-1 `choose_action`."""
    import random
    import torch
    from graph_colouring_with_rl_b200.data import Data
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN
    from graph_colouring_with_rl_b200.envs import GraphColouring, make_target
    from graph_colouring_with_rl_b200.utils import erdos_renyi_adjacency
    rng = np.random.default_rng(1)
    random.seed(1)
    np.random.seed(1)
    torch.manual_seed(1)
    targets = [make_target(erdos_renyi_adjacency(int(n), 0.5, rng), 'g%d' % i, device=dev) for i, n in enumerate(rng.integers(15, 51, size=16))]
    out = {}
    with contextlib.redirect_stdout(sys.stderr):
        for tag in ('env_observations', 'plain_data_int64_edge_index'):
            agent = DQNAgentGN(2, 1, 0, math_mode=args.math)
            env = GraphColouring(targets)
            need = 10 * agent.batch_size
            while agent.memory.current_mem_size < need:
                _, target, cur = env.reset()
                data, gf = env.GenerateData_v1(target, cur)
                done = False
                while not done and agent.memory.current_mem_size < need:
                    v = random.choice(cur['unassigned_vertices'])
                    nxt, reward, done, _ = env.step(v)
                    ndata, ngf = env.GenerateData_v1(target, nxt)
                    if tag == 'plain_data_int64_edge_index':      # what a caller of the reference would hold: CPU tensors
                        plain = lambda d: Data(x=d.x.cpu(), edge_index=d.edge_index.cpu(), edge_attr=d.edge_attr.cpu())
                        agent.remember(plain(data), gf, nxt['assigned_vertices_onehot'], v, reward, plain(ndata), ngf, done)
                    else:
                        agent.remember(data, gf, nxt['assigned_vertices_onehot'], v, reward, ndata, ngf, done)
                    data, gf, cur = ndata, ngf, nxt
            for _ in range(3):
                agent.learn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                agent.learn()
            float(agent.last_loss.item())
            out['learn_ms_' + tag] = (time.perf_counter() - t0) / 20 * 1e3
        # batch-1 act latency (dqn_agent.py:108-140): one 50-vertex graph, mid-episode
        agent = DQNAgentGN(2, 1, 0, test_mode=True, math_mode=args.math)
        t50 = make_target(erdos_renyi_adjacency(50, 0.5, rng), 'n50', device=dev)
        env = GraphColouring([t50])
        target, cur = env.reset(target=t50)
        for v in (3, 17, 29):
            cur, _, _, _ = env.step(v)
        data, gf = env.GenerateData_v1(target, cur)
        for _ in range(5):
            agent.choose_action(data, gf, cur['unassigned_vertices'], test_mode=True)
        t0 = time.perf_counter()
        for _ in range(100):
            agent.choose_action(data, gf, cur['unassigned_vertices'], test_mode=True)
        out['choose_action_ms_n50'] = (time.perf_counter() - t0) / 100 * 1e3
    out['workload'] = ('agent.learn(): 64 transitions of ER(0.5) graphs of 15-50 vertices sampled from 640 stored by agent.remember, host wall '
                       'clock incl. sampling, collate, H2D and the loss read-back, mean of 20; choose_action: one 50-vertex observation, '
                       'wall clock per call incl. the .item() sync, mean of 100')
    return out


def bench_learn_step(learner, batch, timed, args, G):
    """Full DQN learn step on the global minibatch: target forward + fwd/bwd + all-reduce + fused Adam/soft update."""
    learner.learn_step(batch, batch_total=G)
    ms = timed(lambda: learner.learn_step(batch, batch_total=G), max(1, args.steps // 2))
    return {'steps_per_sec': 1e3 / ms, 'transitions_per_sec': G * 1e3 / ms, 'ms_per_step': ms,
            'what': 'training steps/sec of ONE %d-transition minibatch split over the GPUs (strong scaling)' % G}


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


def bench_rollouts(agent, dev, args, rank, world, max_over_ranks, barrier, lib, hbm_peak):
    """Greedy colouring rollouts (BASELINE.json configs[2]): rollout_graphs ER(0.5) graphs of rollout_vertices vertices,
    contiguous shards per rank (strong scaling, no communication), full episodes on the device.  `graphs_per_sec`:
    adjacency resident; `e2e`: host adjacency bytes in -> colours used on the host out (H2D, batching, episodes, D2H)."""
    import ctypes
    import torch
    from graph_colouring_with_rl_b200.envs import BatchedColouringEnv
    from graph_colouring_with_rl_b200.rollouts import greedy_rollouts
    Gr, nr = args.rollout_graphs, args.rollout_vertices
    lo, hi = Gr * rank // world, Gr * (rank + 1) // world
    adjs = er_graphs(Gr, nr, 0.5, seed=0)[lo:hi]
    env = BatchedColouringEnv(list(adjs), device=dev)
    greedy_rollouts(None, agent.gn_behaviour, env=env)             # warm-up sweep
    ms_k = (ctypes.c_double * 8)()
    cnt_k = (ctypes.c_int64 * 8)()
    lib.gcrl_prof_enable(1)
    lib.gcrl_prof_read(ms_k, cnt_k, 8)
    greedy_rollouts(None, agent.gn_behaviour, env=env, graph=False)    # profiled sweep (eager launches: one event pair per kernel)
    torch.cuda.synchronize()
    lib.gcrl_prof_read(ms_k, cnt_k, 8)
    lib.gcrl_prof_enable(0)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    res = greedy_rollouts(None, agent.gn_behaviour, env=env)
    b.record()
    barrier()
    ms = max_over_ranks(a.elapsed_time(b))
    used = res['colours_used'].float().mean().item()
    # e2e: pinned host adjacency -> device env -> episodes -> colours on the host
    flat_adj = torch.from_numpy(np.ascontiguousarray(adjs).reshape(-1)).pin_memory()
    ns = np.full(hi - lo, nr, dtype=np.int64)

    def e2e():
        e = BatchedColouringEnv.from_flat(ns, flat_adj, dev)
        r = greedy_rollouts(None, agent.gn_behaviour, env=e)
        return r['colours_used'].cpu()
    e2e()
    barrier()
    a.record()
    cu = e2e()
    b.record()
    barrier()
    ms_e2e = max_over_ranks(a.elapsed_time(b))
    assert torch.equal(cu, res['colours_used'].cpu())
    # dominant kernel of a rollout step: the edge forward of blocks 2-4 (reads e_{l-1}, writes e_l: 512 B/edge)
    E = int(env.n_edges)
    roof = None
    if cnt_k[1]:
        avg = ms_k[1] / cnt_k[1]
        gbps = 512.0 * E / (avg * 1e-3) / 1e9
        tot = sum(ms_k[i] for i in range(3))
        roof = {'bound': 'hbm', 'kernel': 'k_edge_fwd_wg (blocks 2-4, no stash)', 'achieved': gbps, 'peak': hbm_peak, 'unit': 'GB/s',
                'frac': gbps / hbm_peak, 'avg_ms': avg, 'algorithmic_bytes_per_launch': 512 * E,
                'edge_kernels_share_of_sweep': tot / max(ms, 1e-9) * (1.0 if world == 1 else float('nan')), 'traffic': None}
    return {'graphs_per_sec': Gr / (ms * 1e-3), 'ms_per_sweep': ms, 'steps': res['steps'],
            'workload': '%d ER(0.5) graphs x %d vertices, greedy episodes, default-init weights' % (Gr, nr),
            'mean_colours_used_rank0': used, 'roofline': roof,
            'e2e': {'value': Gr / (ms_e2e * 1e-3), 'unit': 'graphs/s', 'ms_per_sweep': ms_e2e,
                    'h2d_bytes_per_sweep': int(flat_adj.numel()) * (world if world > 1 else 1), 'd2h_bytes_per_sweep': 4 * Gr}}


if __name__ == '__main__':
    if '--cpu-legs' in sys.argv:
        sys.argv.remove('--cpu-legs')
        ap = argparse.ArgumentParser()
        ap.add_argument('--impl')
        ap.add_argument('--vertices', type=int, default=200)
        ap.add_argument('--cpu-sample-graphs', type=int, default=8)
        ap.add_argument('--rollout-vertices', type=int, default=50)
        cpu_legs(ap.parse_args())
    else:
        main()
