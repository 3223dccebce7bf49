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
from .trainer import QLearner
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
                 buffer_size=None, trace=None):
        self.agent = agent
        self.dev = agent.gn_behaviour.device
        self.pool = GraphPool(training_graphs, device=self.dev)
        self.validation_graphs = validation_graphs
        self.n_envs = int(n_envs)
        self.host_rng = (self.n_envs == 1) if host_rng is None else bool(host_rng)
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(int(device_seed))
        self.memory = DeviceReplayBuffer(buffer_size or agent.max_buffer_size, self.pool)
        kw = {} if max_edges_per_micro_batch is None else dict(max_edges_per_micro_batch=max_edges_per_micro_batch)
        # single-GPU loop: without an explicit group the learner stays local even under torchrun (ranks would take
        # data-dependent numbers of learn steps, so a shared all-reduce cadence needs a design of its own)
        self.learner = QLearner(agent, process_group=process_group, local_only=process_group is None, **kw)
        self.validate_every = int(validate_every)
        self.out_dir = out_dir
        self.t_step = 0
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
    def learn(self):
        agent, mem = self.agent, self.memory
        if mem.current_mem_size < 10 * agent.batch_size:
            return None
        if agent.learn_counter == 0:
            print('Starting learning')
            agent.learn_counter += 1
        batch = mem.gather(mem.sample_indices(agent.batch_size))
        loss = self.learner.learn_step(batch)
        self.losses.append(loss)
        return loss

    # ---- validation sweep + best checkpoint (dqn_main.py:126-155, 215-246) ---------------------------------
    def validate(self, ep_ind):
        if self.validation_graphs is None:
            return None
        stats, by_graph = test_using_learned_policy(self.validation_graphs, self.agent, False)
        self.all_validation_stats[ep_ind] = ValidationStats(
            episode_rewards=[g['avg_episode_reward'] for g in by_graph],
            colours_used=[g['avg_colours_used'] for g in by_graph])
        if stats['avg_episode_reward'] >= self.best_validation_reward and self.out_dir is not None:
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
        agent = self.agent
        P = len(self.pool)
        if self.host_rng:
            ids = [random.choice(range(P)) for _ in range(n_episodes)]                 # graph_colouring_env.py:48-50
        else:
            ids = torch.randint(P, (n_episodes,), generator=self.gen, device=self.dev).cpu().numpy().tolist()
        env = BatchedColouringEnv.from_pool(self.pool, ids)
        eps = [epsilon_of(agent, agent.episode_no + e) for e in range(n_episodes)]
        active = np.ones(n_episodes, dtype=bool)
        score = np.zeros(n_episodes, dtype=np.float64)
        seqs = [[] for _ in range(n_episodes)] if self.trace is not None else None
        ep_step = 0
        while active.any():
            a = self._choose(env, ep_step, active, eps)
            cur = env.colour.clone()
            env.step(a)
            if ep_step > 0:                                                             # dqn_main.py:182-184
                self.memory.store_from_env(env, ids, cur, a, mask=active)
            k = int(active.sum())
            # learn once per update_every environment steps (:186-188)
            first = -(-self.t_step // agent.update_every) * agent.update_every
            n_learn = len(range(first, self.t_step + k, agent.update_every))
            self.t_step += k
            for _ in range(n_learn):
                self.learn()
            done_h = env.done.cpu().numpy().astype(bool)                                # the one sync per step
            rew_h = env.reward.cpu().numpy()
            score[active] += rew_h[active]
            if seqs is not None:
                a_h = a.cpu().numpy()
                for e in np.nonzero(active)[0]:
                    seqs[e].append(int(a_h[e]))
            active &= ~done_h
            ep_step += 1
        used = env.colours_used().cpu().numpy()
        for e in range(n_episodes):
            self.episodes_done += 1
            self.training_stats.saved_episodes.append(self.episodes_done)
            self.training_stats.episode_rewards.append(int(score[e]))
            self.training_stats.colours_used.append(int(used[e]))
            agent.episode_no += 1
        if self.trace is not None:
            self.trace.setdefault('graph_ids', []).extend(int(i) for i in ids)
            self.trace.setdefault('actions', []).extend(seqs)
        return ids, used

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
            if self.out_dir is not None:
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
