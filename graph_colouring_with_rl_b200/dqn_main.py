"""The DQN training loop of the reference's dqn_main.py (:158-270) with everything per-step on the
device (SURVEY 8f N1 + N4): batched colouring environments, epsilon-greedy action selection,
device replay ring, learn steps through `QLearner`, periodic validation sweeps as ONE batched
greedy rollout over the validation graphs, best-checkpoint rule and the reference's stat files.

    stats = train(agent, training_graphs, validation_graphs, no_episodes=..., n_envs=...)

`n_envs` episodes run in lock step ("a wave"); finished episodes idle (action -1) until the wave
ends.  What the reference does per environment step is kept per environment step:
  * first vertex of an episode uniformly at random, not stored (:168-171, :182);
  * later vertices epsilon-greedy with the per-episode decay of dqn_agent.py:112-117;
  * `learn()` once per `update_every` environment steps (:186-188), gated on 10 x batch_size
    stored transitions (dqn_agent.py:149);
  * validation every `validate_every` episodes and before the first one (:126, :215), checkpoint
    kept when avg reward >= best so far, previous best unlinked (:145-155, :236-246).
With `n_envs=1, host_rng=True` the loop consumes `random` / `np.random` in exactly the
reference's order (env.reset's graph choice, first vertex, choose_action's random.random() and
random.choice, the sampler's np.random.choice), so a seeded run takes the reference's branches;
larger waves draw exploration from a device generator instead.

Multi-GPU (pass `process_group`, or `distributed=True` for the default group; one process per GPU): the WAVE is
sharded by graph -- rank r runs its contiguous share of the wave's episodes in its own batched environment and
keeps the transitions it produced in its own replay ring -- and every learn step is ONE global minibatch of
`batch_size` transitions, `batch_size / world` sampled from each rank's ring, with the single gradient all-reduce of
`QLearner`.  The learn cadence is a function of the GLOBAL environment-step count (one tiny all-reduce of
[active episodes, stored transitions] per wave step, next to the host sync the step already has), so every rank
posts the same number of collectives, finished ranks keep stepping (with no-op actions) until the whole wave is
done, and the replicas stay bit-identical.  Validation sweeps are sharded by graph and all-gathered.
"""
import os
import random
from collections import namedtuple

import numpy as np
import torch

from . import ops
from .envs import BatchedColouringEnv
from .replay import DeviceReplayBuffer, GraphPool
from .test_policy_functions import test_using_learned_policy
from .trainer import QLearner, TransitionBatch, shard_range
from .utils import save_benchmark_stats_to_file, save_training_stats_to_file, save_validation_stats_to_file

TrainingStats = namedtuple("Stats", ["saved_episodes", "episode_rewards", "colours_used"])
ValidationStats = namedtuple("Stats", ["episode_rewards", "colours_used"])


def epsilon_of(agent, episode_no):
    """dqn_agent.py:112-117."""
    epsilon_decay = (agent.eps_end / agent.eps_start) ** (1 / agent.no_episodes)
    return agent.eps_start * (epsilon_decay ** episode_no)


class DeviceTrainer(object):
    def __init__(self, agent, training_graphs, validation_graphs=None, n_envs=1, host_rng=None, device_seed=0,
                 validate_every=500, out_dir=None, max_edges_per_micro_batch=None, process_group=None,
                 buffer_size=None, trace=None, distributed=False):
        self.agent = agent
        self.dev = agent.gn_behaviour.device
        self.pool = GraphPool(training_graphs, device=self.dev)
        self.validation_graphs = validation_graphs
        self.n_envs = int(n_envs)
        self.memory = DeviceReplayBuffer(buffer_size or agent.max_buffer_size, self.pool)
        kw = {} if max_edges_per_micro_batch is None else dict(max_edges_per_micro_batch=max_edges_per_micro_batch)
        # without an explicit group (or distributed=True) the learner stays local even under torchrun: a trainer that only
        # some ranks run must not enter collectives the other ranks never post
        multi = process_group is not None or distributed
        self.learner = QLearner(agent, process_group=process_group, local_only=not multi, **kw)
        self.group = process_group
        self.world = self.learner.world
        self.rank = torch.distributed.get_rank(process_group) if self.world > 1 else 0
        self.host_rng = (self.n_envs == 1 and self.world == 1) if host_rng is None else bool(host_rng)
        if self.host_rng and self.world > 1:
            raise ValueError('host_rng reproduces the single-process reference run; it has no multi-GPU meaning')
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(int(device_seed) + 7919 * self.rank)          # ranks explore independently
        self.validate_every = int(validate_every)
        self.out_dir = out_dir
        self.t_step = 0
        self._stored_g = 0
        self.episodes_done = 0
        self.training_stats = TrainingStats([], [], [])
        self.all_validation_stats = {}
        self.best_validation_reward = -np.inf
        self.best_validation_filepath = None
        self.losses = []                 # device scalars, one per learn step
        self.trace = trace               # optional dict of lists: per-episode graph ids / action sequences
        if out_dir is not None:
            self.stats_dir = os.path.join(out_dir, 'training_analysis')
            self.params_dir = os.path.join(out_dir, 'trained_params')
            os.makedirs(self.stats_dir, exist_ok=True)
            os.makedirs(self.params_dir, exist_ok=True)

    # ---- learn (dqn_agent.py:147-223 over the device ring) ---------------------------------------------
    def _stored_global(self):
        return self._stored_g if self.world > 1 else self.memory.current_mem_size

    def learn(self):
        agent, mem = self.agent, self.memory
        if self._stored_global() < 10 * agent.batch_size:
            return None
        if agent.learn_counter == 0:
            print('Starting learning')
            agent.learn_counter += 1
        if self.world == 1:
            batch = mem.gather(mem.sample_indices(agent.batch_size))
            loss = self.learner.learn_step(batch)
        else:
            lo, hi = shard_range(agent.batch_size, self.rank, self.world)
            k = (hi - lo) if mem.current_mem_size > 0 else 0      # a rank whose ring is still empty joins with no graphs
            batch = mem.gather(mem.sample_indices(k)) if k else TransitionBatch.empty(self.dev)
            loss = self.learner.learn_step(batch)                  # global mean: the graph count rides in the all-reduce
        self.losses.append(loss)
        return loss

    # ---- validation sweep + best checkpoint (dqn_main.py:126-155, 215-246) ---------------------------------
    def validate(self, ep_ind):
        if self.validation_graphs is None:
            return None
        if self.world > 1:
            lo, hi = shard_range(len(self.validation_graphs), self.rank, self.world)
            _, mine = test_using_learned_policy(self.validation_graphs[lo:hi], self.agent, False) if hi > lo else (None, [])
            parts = [None] * self.world
            torch.distributed.all_gather_object(parts, mine, group=self.group)
            by_graph = [g for p in parts for g in p]
            rewards = np.asarray([g['avg_episode_reward'] for g in by_graph])
            used = np.asarray([g['avg_colours_used'] for g in by_graph])
            stats = {"avg_max_episode_reward": np.mean(rewards), "avg_episode_reward": np.mean(rewards),
                     "episode_reward_std": np.std(rewards), "avg_min_colours_used": np.mean(used),
                     "avg_colours_used": np.mean(used), "colours_used_std": np.std(used)}
        else:
            stats, by_graph = test_using_learned_policy(self.validation_graphs, self.agent, False)
        self.all_validation_stats[ep_ind] = ValidationStats(
            episode_rewards=[g['avg_episode_reward'] for g in by_graph],
            colours_used=[g['avg_colours_used'] for g in by_graph])
        if stats['avg_episode_reward'] >= self.best_validation_reward and self.out_dir is not None and self.rank == 0:
            if self.best_validation_filepath is not None:
                os.unlink(self.best_validation_filepath)
            self.best_validation_reward = stats['avg_episode_reward']
            root = os.path.join(self.params_dir, 'best_policy_episode_' + str(ep_ind))
            self.best_validation_filepath = self.agent.save_models(root)
        return stats

    def save_stats(self, ep_ind, benchmark_stats=None):
        """dqn_main.py:254-270: one file per kind, older ones removed."""
        for f in os.listdir(self.stats_dir):
            fp = os.path.join(self.stats_dir, f)
            if os.path.isfile(fp):
                os.unlink(fp)
        if benchmark_stats is not None:
            save_benchmark_stats_to_file(benchmark_stats, os.path.join(self.stats_dir, 'benchmark_stats'))
        save_training_stats_to_file(self.training_stats,
                                    os.path.join(self.stats_dir, 'training_stats_episode_' + str(ep_ind) + '.txt'))
        save_validation_stats_to_file(self.all_validation_stats,
                                      os.path.join(self.stats_dir, 'validation_stats_episode_' + str(ep_ind) + '.txt'))

    # ---- action selection ------------------------------------------------------------------------------------
    def _random_vertices(self, env):
        """Uniformly random uncoloured vertex per graph from the device generator."""
        r = torch.rand(env.n_vertices, generator=self.gen, device=self.dev)
        return ops.score_argmax(r, env.colour, env.node_off)

    def _greedy(self, env):
        net = self.agent.gn_behaviour
        packed, eattr = env.packed
        _, q, _ = ops.gn_forward(net.flat_params(), net.dims, packed, env.x, eattr, math_mode=net.math_mode)
        return ops.score_argmax(q, env.colour, env.node_off)

    def _choose(self, env, ep_step, active, eps):
        """-> int32 [n_envs] device actions (-1 for finished episodes)."""
        G = env.n_graphs
        if self.host_rng:
            # reference order, one environment after the other
            col = env.colour.cpu().numpy()
            off = env.node_off.cpu().numpy()
            a = np.full(G, -1, dtype=np.int32)
            greedy = None
            for e in range(G):
                if not active[e]:
                    continue
                un = np.nonzero(col[off[e]:off[e + 1]] < 0)[0].tolist()
                if ep_step == 0:
                    a[e] = random.choice(un)                                         # dqn_main.py:171
                elif random.random() > eps[e]:                                       # dqn_agent.py:119
                    if greedy is None:
                        greedy = self._greedy(env).cpu().numpy()
                    a[e] = greedy[e]
                else:
                    a[e] = random.choice(un)                                         # :137
            return torch.from_numpy(a).to(self.dev)
        act_t = torch.from_numpy(active).to(self.dev)
        a_rand = self._random_vertices(env)
        if ep_step == 0:
            a = a_rand
        else:
            u = torch.rand(G, generator=self.gen, device=self.dev)
            eps_t = torch.from_numpy(np.asarray(eps, dtype=np.float32)).to(self.dev)
            a = torch.where(u > eps_t, self._greedy(env), a_rand)
        return torch.where(act_t, a, torch.full_like(a, -1))

    # ---- one wave of n_envs episodes ---------------------------------------------------------------------------
    @torch.no_grad()
    def run_wave(self, n_episodes):
        """n_episodes: episodes of the whole wave (all ranks); this rank runs its contiguous share."""
        agent = self.agent
        P = len(self.pool)
        e_lo, e_hi = shard_range(n_episodes, self.rank, self.world)
        n_local = e_hi - e_lo
        if self.host_rng:
            ids = [random.choice(range(P)) for _ in range(n_local)]                    # graph_colouring_env.py:48-50
        else:
            ids = torch.randint(P, (max(n_local, 1),), generator=self.gen, device=self.dev).cpu().numpy().tolist()[:n_local]
        env = BatchedColouringEnv.from_pool(self.pool, ids) if n_local else None
        eps = [epsilon_of(agent, agent.episode_no + e_lo + e) for e in range(n_local)]
        active = np.ones(n_local, dtype=bool)
        score = np.zeros(n_local, dtype=np.float64)
        seqs = [[] for _ in range(n_local)] if self.trace is not None else None
        ep_step = 0
        counts = torch.zeros(2, dtype=torch.int64, device=self.dev) if self.world > 1 else None
        while True:
            k = int(active.sum())
            if self.world > 1:
                # the learn cadence and the end of the wave are GLOBAL: [active episodes, stored transitions] summed over
                # the ranks, BEFORE this step's transitions are stored (what the single-process loop sees at this point)
                counts[0], counts[1] = k, self.memory.current_mem_size
                torch.distributed.all_reduce(counts, group=self.group)
                k_g, self._stored_g = (int(v) for v in counts.tolist())
            else:
                k_g = k
            if k_g == 0:
                break
            if k:
                a = self._choose(env, ep_step, active, eps)
                cur = env.colour.clone()
                env.step(a)
                if ep_step > 0:                                                         # dqn_main.py:182-184
                    self.memory.store_from_env(env, ids, cur, a, mask=active)
            if self.world > 1 and ep_step > 0:
                self._stored_g += k_g                # every active episode of every rank stored one transition
            # learn once per update_every environment steps (:186-188), counted over all ranks
            first = -(-self.t_step // agent.update_every) * agent.update_every
            n_learn = len(range(first, self.t_step + k_g, agent.update_every))
            self.t_step += k_g
            for _ in range(n_learn):
                self.learn()
            if k:
                done_h = env.done.cpu().numpy().astype(bool)                            # the one sync per step
                rew_h = env.reward.cpu().numpy()
                score[active] += rew_h[active]
                if seqs is not None:
                    a_h = a.cpu().numpy()
                    for e in np.nonzero(active)[0]:
                        seqs[e].append(int(a_h[e]))
                active &= ~done_h
            ep_step += 1
        used = env.colours_used().cpu().numpy() if n_local else np.zeros(0, dtype=np.int64)
        mine = ([int(i) for i in ids], [int(v) for v in score], [int(u) for u in used])
        if self.world > 1:
            parts = [None] * self.world
            torch.distributed.all_gather_object(parts, mine, group=self.group)
        else:
            parts = [mine]
        all_ids, all_used = [], []
        for p_ids, p_score, p_used in parts:
            for g, sc, u in zip(p_ids, p_score, p_used):
                self.episodes_done += 1
                self.training_stats.saved_episodes.append(self.episodes_done)
                self.training_stats.episode_rewards.append(sc)
                self.training_stats.colours_used.append(u)
                agent.episode_no += 1
                all_ids.append(g)
                all_used.append(u)
        if self.trace is not None:
            self.trace.setdefault('graph_ids', []).extend(int(i) for i in ids)
            self.trace.setdefault('actions', []).extend(seqs)
        return all_ids, np.asarray(all_used)

    def run(self, no_episodes=None, log_every=100, benchmark_stats=None, verbose=True):
        agent = self.agent
        no_episodes = agent.no_episodes if no_episodes is None else int(no_episodes)
        if self.episodes_done == 0:
            stats = self.validate(0)                                                    # dqn_main.py:126
            if verbose and stats is not None:
                print('Average reward: %.2f' % stats['avg_episode_reward'],
                      '| Average colours used: %.2f' % stats['avg_colours_used'])
        while self.episodes_done < no_episodes:
            before = self.episodes_done
            # a wave never straddles a validation point, so the sweep sees the policy of that episode count
            nxt = min(no_episodes, (before // self.validate_every + 1) * self.validate_every)
            self.run_wave(min(self.n_envs, nxt - before))
            ep = self.episodes_done
            if verbose and ep // log_every != before // log_every:
                print('episode: ', ep, '| 100 game average reward: %.2f' % np.mean(self.training_stats.episode_rewards[-100:]),
                      '| 100 game average colours_used: %.2f' % np.mean(self.training_stats.colours_used[-100:]))
            if ep % self.validate_every == 0:
                stats = self.validate(ep)
                if verbose and stats is not None:
                    print('Current performance on validation graphs:')
                    print('Average reward: %.2f' % stats['avg_episode_reward'],
                          '| Average colours used: %.2f' % stats['avg_colours_used'])
            if self.out_dir is not None and self.rank == 0:
                if ep == no_episodes:
                    agent.save_models(os.path.join(self.params_dir, 'final_policy_episode_' + str(ep)))
                if ep // 1000 != before // 1000 or ep == no_episodes:
                    self.save_stats(ep, benchmark_stats)
        return self.training_stats


def train(agent, training_graphs, validation_graphs=None, no_episodes=None, **kw):
    """Convenience wrapper: build a DeviceTrainer and run it.  Returns the trainer."""
    run_kw = {k: kw.pop(k) for k in ('log_every', 'benchmark_stats', 'verbose') if k in kw}
    t = DeviceTrainer(agent, training_graphs, validation_graphs, **kw)
    t.run(no_episodes, **run_kw)
    return t
