"""Host-side helpers either side of the hot path (SURVEY 8f N3 + N4), with the reference's names
and return values: stat files (utils.py:58-125), DIMACS `.col` reader (utils.py:243-282,
447-466), the complete-graph edge encoding (utils.py:296-323), queen graphs
(construct_queen_graph.py:3-67), Leighton graphs (construct_leighton_graph.py).  Pure numpy / Python: nothing here touches the device, and
the graphs these functions describe are handed to the kernels as dense adjacency bytes
(`adjacency_from_*`), never as the O(n^2)-tuple Python structures the reference builds.
"""
import random
from collections import namedtuple

import numpy as np


# ---- stat files (same text format as utils.py:58-125) -----------------------------------------------------
def save_training_stats_to_file(stats, filename):
    with open(filename, 'w') as file:
        file.write(','.join(np.array(stats.saved_episodes).astype(str)))
        file.write('\n')
        file.write(','.join(np.array(stats.episode_rewards).astype(str)))
        file.write('\n')
        file.write(','.join(np.array(stats.colours_used).astype(str)))


def save_validation_stats_to_file(validation_stats, filename):
    with open(filename, 'w') as file:
        for episode, stats in validation_stats.items():
            file.write(str(episode) + '\n')
            file.write(','.join(np.array(stats.episode_rewards).astype(str)) + '\n')
            file.write(','.join(np.array(stats.colours_used).astype(str)) + '\n')


def save_benchmark_stats_to_file(benchmark_stats, filename):
    with open(filename, 'w') as file:
        file.write(','.join(np.array(benchmark_stats.episode_rewards).astype(str)))
        file.write('\n')
        file.write(','.join(np.array(benchmark_stats.colours_used).astype(str)))


def _floats(line):
    return [float(item) for item in line.split(',')]


def load_stats_from_file(filepath):
    with open(filepath) as file:
        lines = file.readlines()
    return namedtuple("Stats", ["saved_episodes", "episode_rewards", "colours_used"])(
        saved_episodes=_floats(lines[0]), episode_rewards=_floats(lines[1]), colours_used=_floats(lines[2]))


def load_validation_stats_from_file(filepath):
    validation_stats = {}
    with open(filepath) as file:
        lines = file.readlines()
    if not len(lines) % 3 == 0:
        raise Exception('No of lines in stats file should be multiple of 3')
    for ind in range(len(lines) // 3):
        validation_stats[int(lines[3 * ind])] = namedtuple("Stats", ["episode_rewards", "colours_used"])(
            episode_rewards=_floats(lines[3 * ind + 1]), colours_used=_floats(lines[3 * ind + 2]))
    return validation_stats


def load_benchmark_stats_from_file(filepath):
    with open(filepath) as file:
        lines = file.readlines()
    return namedtuple("Stats", ["episode_rewards", "colours_used"])(
        episode_rewards=_floats(lines[0]), colours_used=_floats(lines[1]))


# ---- complete-graph edge encoding ---------------------------------------------------------------------------
def complete_edge_arrays(n):
    """The canonical directed pair order (0,1),(1,0),(0,2),(2,0),... of utils.py:304-310 as an
    int64 [n(n-1), 2] array (vectorised; the reference builds it tuple by tuple)."""
    a, b = np.triu_indices(n, 1)
    el = np.empty((a.size * 2, 2), dtype=np.int64)
    el[0::2, 0], el[0::2, 1] = a, b
    el[1::2, 0], el[1::2, 1] = b, a
    return el


def edge_index_of(u, v, n):
    """Position of the ordered pair (u, v), u != v, in the canonical order (what the reference
    looks up in edge_ind_dict[u][v])."""
    u, v = np.asarray(u, dtype=np.int64), np.asarray(v, dtype=np.int64)
    lo, hi = np.minimum(u, v), np.maximum(u, v)
    base = lo * (2 * n - lo - 1) // 2 + (hi - lo - 1)      # rank of {lo, hi} among unordered pairs
    return 2 * base + (u > v)


def _edge_ind_dict(el):
    d = {}
    for k, (u, v) in enumerate(el.tolist()):
        d.setdefault(u, {})[v] = k
    return d


def encode_adjacency(adj):
    """Dense 0/1 adjacency -> (edge_list, edge_attr, edge_ind_dict) in the reference's encoding
    (edge_attr -1 = edge, 0 = non-edge)."""
    adj = np.asarray(adj)
    n = adj.shape[0]
    el = complete_edge_arrays(n)
    ea = -(adj[el[:, 0], el[:, 1]] != 0).astype(np.int64)
    ind = _edge_ind_dict(el)
    for v in range(n):
        ind.setdefault(v, {})
    return [tuple(e) for e in el.tolist()], ea.tolist(), ind


def adjacency_from_edges(n, edges):
    """Undirected edge pairs (0-based) -> symmetric uint8 [n,n]; self loops dropped."""
    adj = np.zeros((n, n), dtype=np.uint8)
    e = np.asarray(list(edges), dtype=np.int64).reshape(-1, 2)
    e = e[e[:, 0] != e[:, 1]]
    adj[e[:, 0], e[:, 1]] = 1
    adj[e[:, 1], e[:, 0]] = 1
    return adj


def convert_networkx_to_graph(vertices, edges):
    """utils.py:296-323: -> (edge_list tuple, edge_attr list, edge_ind_dict, target_edge_indices list
    in the order of `edges`, two entries per edge)."""
    nodes = list(vertices)
    pos = {v: i for i, v in enumerate(nodes)}
    n = len(nodes)
    el = complete_edge_arrays(n)
    edge_list = tuple((nodes[u], nodes[v]) for u, v in el.tolist())
    ind = {v: {} for v in nodes}
    for k, (u, v) in enumerate(el.tolist()):
        ind[nodes[u]][nodes[v]] = k
    attr = [0] * len(edge_list)
    indices = []
    for v1, v2 in list(edges):
        for a, b in ((v1, v2), (v2, v1)):
            k = int(edge_index_of(pos[a], pos[b], n))
            attr[k] = -1
            indices.append(k)
    return edge_list, attr, ind, indices


# ---- DIMACS .col files ------------------------------------------------------------------------------------------
def _parse_dimacs(full_graph_filepath):
    n, edges = None, []
    with open(full_graph_filepath, 'r') as f:
        for line in f.read().splitlines():
            x = line.split(' ')
            if x[0] == 'p':
                n = int(x[2])
            elif x[0] == 'e':
                edges.append((int(x[1]) - 1, int(x[2]) - 1))
    if n is None:
        raise ValueError('%s: no "p edge <vertices> <edges>" line' % full_graph_filepath)
    return n, edges


def import_dimacs_adjacency(full_graph_filepath):
    """DIMACS edge file -> dense symmetric uint8 adjacency (what the device kernels take)."""
    n, edges = _parse_dimacs(full_graph_filepath)
    return adjacency_from_edges(n, edges)


def import_dimacs_graph(full_graph_filepath):
    """utils.py:243-282: -> (no_vertices, edge_list tuple, edge_attr tuple, edge_ind_dict,
    tuple(set(target_edge_indices)))."""
    n, edges = _parse_dimacs(full_graph_filepath)
    el = complete_edge_arrays(n)
    attr = np.zeros(el.shape[0], dtype=np.int64)
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    e = e[e[:, 0] != e[:, 1]]
    k = np.concatenate([edge_index_of(e[:, 0], e[:, 1], n), edge_index_of(e[:, 1], e[:, 0], n)])
    attr[k] = -1
    ind = _edge_ind_dict(el)
    for v in range(n):
        ind.setdefault(v, {})
    return (n, tuple(tuple(p) for p in el.tolist()), tuple(attr.tolist()), ind,
            tuple(set(int(i) for i in k.tolist())))


def import_dimacs_graph_as_networkx(full_graph_filepath):
    """utils.py:447-466 (only edges with vertex1 < vertex2 are added, as there)."""
    import networkx as nx
    n, edges = _parse_dimacs(full_graph_filepath)
    G = nx.Graph()
    G.add_nodes_from(range(n))
    G.add_edges_from((u, v) for u, v in edges if u < v)
    return G


# ---- definitional graph families ---------------------------------------------------------------------------------
def queen_adjacency(no_rows, no_cols):
    """Queen graph on a no_rows x no_cols board, vertex v at (v // no_cols, v % no_cols)
    (construct_queen_graph.py:36-63: same row, same column or same diagonal)."""
    v = np.arange(no_rows * no_cols)
    r, c = v // no_cols, v % no_cols
    dr, dc = r[:, None] - r[None, :], c[:, None] - c[None, :]
    adj = (dr == 0) | (dc == 0) | (np.abs(dr) == np.abs(dc))
    np.fill_diagonal(adj, False)
    return adj.astype(np.uint8)


def generate_queen_graph(n, k):
    """construct_queen_graph.py:3-67: board k x n/k or n/k x k (one random.random() draw), ->
    (edge_list, edge_attr, edge_ind_dict, target_edge_indices)."""
    if not n % k == 0:
        raise Exception('n must be a multiple of k for a Queen graphs')
    if random.random() > 0.5:
        no_rows, no_cols = k, n // k
    else:
        no_cols, no_rows = k, n // k
    adj = queen_adjacency(no_rows, no_cols)
    edge_list, edge_attr, ind = encode_adjacency(adj)
    indices = [i for i, a in enumerate(edge_attr) if a == -1]
    return edge_list, edge_attr, ind, indices


def mycielski_adjacency(k):
    """Mycielski graph M_k (DIMACS myciel<k-1>: myciel3 = M_4, ...; M_2 = K_2): chromatic number k,
    triangle free.  Same vertex numbering as networkx.mycielski_graph."""
    adj = np.zeros((1, 1), dtype=np.uint8) if k == 1 else np.array([[0, 1], [1, 0]], dtype=np.uint8)
    for _ in range(max(0, k - 2)):
        n = adj.shape[0]
        out = np.zeros((2 * n + 1, 2 * n + 1), dtype=np.uint8)
        out[:n, :n] = adj
        out[:n, n:2 * n] = adj          # v_i -- u_j for every edge v_i v_j
        out[n:2 * n, :n] = adj.T
        out[n:2 * n, 2 * n] = 1         # every u_i -- w
        out[2 * n, n:2 * n] = 1
        adj = out
    return adj


def erdos_renyi_adjacency(n, p, rng):
    """G(n, p) from a numpy Generator: upper-triangular Bernoulli mirrored."""
    up = np.triu(rng.random((n, n)) < p, 1)
    return (up | up.T).astype(np.uint8)


# ---- Leighton graphs (construct_leighton_graph.py) -----------------------------------------------------------------
def _primes_below(limit):
    sieve = np.ones(limit, dtype=bool)
    sieve[:2] = False
    for p in range(2, int(limit ** 0.5) + 1):
        if sieve[p]:
            sieve[p * p::p] = False
    return [int(p) for p in np.nonzero(sieve)[0]]


_PRIMES = _primes_below(7920)          # the reference's hard-coded table: the first 1000 primes, 2..7919


def get_prime_factors(n):
    """construct_leighton_graph.py:get_prime_factors (trial division, with multiplicity for 2 only as there)."""
    factors = []
    i = 2
    while n % i == 0:
        n //= i
        factors.append(i)
    i += 1
    while i * i <= n:
        if n % i == 0:
            n //= i
            factors.append(i)
        else:
            i += 2
    if n > 1:
        factors.append(n)
    return factors


def generate_leighton_parameters(n, k):
    """(m, c, a, b) of Leighton's generator with the reference's draws in the reference's order
    (generate_m, generate_c, generate_a, generate_b): m >> n with hcf(m, n) = k, hcf(c, m) = 1, every prime of m
    (and 4, if 4 | m) divides a - 1, b[i] = number of (k - i)-cliques."""
    n_primes = set(get_prime_factors(n))
    m_factors = get_prime_factors(k)
    remaining = [p for p in _PRIMES if p not in n_primes]
    m = k
    while m < k * n ** 2:
        p = random.choice(remaining)
        m_factors.append(p)
        m *= p
    remaining_c = [p for p in _PRIMES if p not in set(m_factors)]
    c = 1
    while c < 10 * k * n:
        c *= random.choice(remaining_c)
    a_minus_1 = 4 if m % 4 == 0 else 1
    for p in set(m_factors):
        a_minus_1 *= p
    while a_minus_1 < 10 * k * n:
        a_minus_1 *= random.choice(_PRIMES)
    b = [random.randrange(1, int(n / k) + 1)] + [random.randrange(int(n / (k - i) + 1)) for i in range(1, k - 1)]
    return m, c, a_minus_1 + 1, b


def leighton_edges(n, k, m, c, a, b, X_0):
    """generate_Xi / generate_Yi / generate_edges: the linear congruential sequence X_i = (a X_{i-1} + c) mod m, Y_i = X_i
    mod n, consecutive runs of Y as cliques of size k, k-1, ... (b[i] of each).  Returns the list of (min, max) pairs
    in generation order (duplicates and self pairs included, as in the reference before its set())."""
    sizes = [len(b) - i + 1 for i in range(len(b))]
    length = sum(nc * s for nc, s in zip(b, sizes)) + 1
    xs = [X_0]
    for _ in range(1, length):
        xs.append((a * xs[-1] + c) % m)
    ys = [x % n for x in xs]
    edges, pos = [], 1
    for i, nb in enumerate(b):
        size = k - i
        for _ in range(nb):
            vs = ys[pos:pos + size]
            pos += size
            for j, v1 in enumerate(vs):
                for v2 in vs[j + 1:]:
                    edges.append((min(v1, v2), max(v1, v2)))
    return edges


def generate_leighton_graph(n, k, X_0=None):
    """construct_leighton_graph.py:generate_leighton_graph -> (edge_list, edge_attr, edge_ind_dict,
    target_edge_indices); same random draws, same set-iteration order of target_edge_indices."""
    m, c, a, b = generate_leighton_parameters(n, k)
    if X_0 is None:
        X_0 = random.randrange(m)
    graph_edges = set(leighton_edges(n, k, m, c, a, b, X_0))
    el = complete_edge_arrays(n)
    attr = [0] * el.shape[0]
    ind = _edge_ind_dict(el)
    for v in range(n):
        ind.setdefault(v, {})
    indices = []
    for v1, v2 in graph_edges:
        # a clique that draws the same vertex twice yields a self pair; the reference's dict lookup raises KeyError there
        i1, i2 = ind[int(v1)][int(v2)], ind[int(v2)][int(v1)]
        attr[i1] = -1
        indices.append(i1)
        attr[i2] = -1
        indices.append(i2)
    return [tuple(e) for e in el.tolist()], attr, ind, indices


def leighton_adjacency(n, k, X_0=None):
    """Dense adjacency of a Leighton graph (self pairs dropped instead of raising)."""
    m, c, a, b = generate_leighton_parameters(n, k)
    if X_0 is None:
        X_0 = random.randrange(m)
    return adjacency_from_edges(n, [e for e in leighton_edges(n, k, m, c, a, b, X_0) if e[0] != e[1]])
