"""Drop-in for the reference's test_policy_functions.py: `test_using_learned_policy(graphs,
agent, stochastic)` with the same return value `(stats_all_graphs, stats_by_graph)`, evaluated
as ONE batched device rollout over all graphs (rollouts.greedy_rollouts) instead of a Python
episode loop per graph.  Reference: test_policy_functions.py:6-103.
"""
import numpy as np

from .rollouts import greedy_rollouts

__test__ = False     # not a pytest module (the reference's file name collides with pytest's pattern)


def test_using_learned_policy(graphs, agent, stochastic):
    if stochastic:
        # the reference's stochastic branch calls agent.choose_action_stochastic_2, which does not
        # exist anywhere in the reference (test_policy_functions.py:54)
        raise AttributeError("'DQNAgentGN' object has no attribute 'choose_action_stochastic_2'")
    res = greedy_rollouts(graphs, agent.gn_behaviour)
    used = res['colours_used'].cpu().numpy().astype(np.int64)
    stats_by_graph = []
    for target, c in zip(graphs, used):
        ep = -int(c)                       # sum of per-step rewards -(new colours) = -colours used
        stats_by_graph.append({
            "name": target['name'] if isinstance(target, dict) and 'name' in target else None,
            "avg_episode_reward": float(ep), "episode_reward_std": 0.0, "max_episode_reward": ep,
            "avg_colours_used": float(c), "colours_used_std": 0.0, "min_colours_used": int(c),
        })
    rewards = -used
    stats_all_graphs = {
        "avg_max_episode_reward": np.mean(rewards), "avg_episode_reward": np.mean(rewards),
        "episode_reward_std": np.std(rewards), "avg_min_colours_used": np.mean(used),
        "avg_colours_used": np.mean(used), "colours_used_std": np.std(used),
    }
    return stats_all_graphs, stats_by_graph
