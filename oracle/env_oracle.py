"""TEST INFRASTRUCTURE ONLY -- integer restatement (numpy + small Python loops) of the
reference's colouring environment and rollout harness.

Pinned: replaying the DSATUR `vertex_order` stored for each of the 100 graphs of the
reference's datasets/validation_dataset.pickle through `ColouringEnvOracle.step` ends `done`
with exactly the stored `colours_used` (tests/golden/validation_graphs.npz, written by
oracle/make_golden.py), and step-by-step colours/rewards/done equal the unmodified
reference env on the same action sequences (tests/golden/env_traces.npz).

Reference lines followed:
  graph_colouring_env.py:64-90   step            :102-109 GenerateRewards
  graph_colouring_env.py:144-182 colour_vertex   :185-191 colour_leaf_nodes
  utils.py:285-293 get_vertex_colour (first fit)
  utils.py:522-533, 559-565 <=3-hop distances and per-distance neighbour counts
  test_policy_functions.py:46-47 first-vertex rule
"""
import numpy as np

MAX_RADIUS_FOR_NEIGHBOURS = 3  # utils.py:19


def first_fit(neighbour_colours):
    """utils.py:285-293 with colour names 0,1,2,... (graph_colouring_env.py:53)."""
    c = 0
    while c in neighbour_colours:
        c += 1
    return c


def neighbour_dist_counts(adj):
    """utils.py:522-533 + 559-565: number of vertices at exact shortest-path distance
    1, 2, 3 from each vertex (the reference's relaxation yields true distances capped at 3)."""
    n = adj.shape[0]
    a = adj.astype(bool)
    reach = np.eye(n, dtype=bool)
    frontier = np.eye(n, dtype=bool)
    counts = np.zeros((n, MAX_RADIUS_FOR_NEIGHBOURS), dtype=np.float64)
    for d in range(MAX_RADIUS_FOR_NEIGHBOURS):
        nxt = (frontier.astype(np.int64) @ a.astype(np.int64)) > 0
        nxt &= ~reach
        counts[:, d] = nxt.sum(1)
        reach |= nxt
        frontier = nxt
    return counts


def first_vertex(adj):
    """test_policy_functions.py:46-47: argmax_v sum_i counts[v][i] * n**(2-i), first max."""
    n = adj.shape[0]
    counts = neighbour_dist_counts(adj)
    mult = np.array([n ** (MAX_RADIUS_FOR_NEIGHBOURS - 1 - i) for i in range(MAX_RADIUS_FOR_NEIGHBOURS)])
    return int(np.argmax(np.sum(counts * mult, axis=1)))


class ColouringEnvOracle(object):
    """Minimal state equivalent to the reference's `current_graph` dict:
    colour[v] (-1 = uncoloured) and max colour index used."""

    def __init__(self, adj):
        self.adj = np.asarray(adj).astype(bool)
        self.n = self.adj.shape[0]
        self.reset()

    def reset(self):
        self.colour = -np.ones(self.n, dtype=np.int64)
        self.max_colour = -1
        return self.colour.copy()

    @property
    def no_colours_used(self):
        return self.max_colour + 1                       # graph_colouring_env.py:158-159

    @property
    def unassigned_vertices(self):
        return [int(v) for v in np.nonzero(self.colour < 0)[0]]

    @property
    def assigned_onehot(self):
        return (self.colour >= 0).astype(np.int64)

    def _neighbour_colours(self, v):
        nb = self.colour[self.adj[v]]
        return set(int(c) for c in nb[nb >= 0])

    def _colour_vertex(self, v):
        c = first_fit(self._neighbour_colours(v))        # :70, utils.py:285-293
        self.colour[v] = c                               # :145
        self.max_colour = max(self.max_colour, c)        # :158
        return c

    def step(self, v):
        """graph_colouring_env.py:64-90 -> (colour vector, reward, done)."""
        if not (0 <= v < self.n):
            raise Exception(str(v) + ' is not a vertex name.')       # :72-73
        prev = self.no_colours_used
        self._colour_vertex(int(v))                                  # :78
        for u in self.unassigned_vertices:                           # :185-191 (snapshot, ascending)
            if not np.any(self.adj[u] & (self.colour < 0)):
                self._colour_vertex(u)
        done = bool(np.all(self.colour >= 0))                        # :82
        reward = -(self.no_colours_used - prev)                      # :107
        return self.colour.copy(), reward, done


def replay_order(adj, order):
    env = ColouringEnvOracle(adj)
    done = False
    for v in order:
        _, _, done = env.step(int(v))
    return env.no_colours_used, done, env.colour.copy()


def greedy_rollout(adj, q_fn, edge_list=None):
    """test_policy_functions.py:36-73: first vertex by the degree rule, then argmax-Q over
    unassigned vertices until done.  q_fn(colour_vector) -> q [n].  Returns
    (colours_used, vertex_sequence, final colour vector, per-step top-2 Q gaps)."""
    env = ColouringEnvOracle(adj)
    seq, gaps = [], []
    v = first_vertex(adj)
    done = False
    while True:
        seq.append(v)
        _, _, done = env.step(v)
        if done:
            break
        q = np.asarray(q_fn(env.colour.copy()), dtype=np.float32).reshape(-1)
        un = env.unassigned_vertices
        qs = q[un]
        v = int(un[int(np.argmax(qs))])
        srt = np.sort(qs)
        gaps.append(float(srt[-1] - srt[-2]) if len(qs) > 1 else float('inf'))
    return env.no_colours_used, seq, env.colour.copy(), gaps


def er_graph(n, p, rng):
    """Synthetic Erdos-Renyi adjacency: upper-triangular Bernoulli mirrored."""
    u = rng.random((n, n)) < p
    a = np.triu(u, 1)
    return (a | a.T).astype(np.uint8)
