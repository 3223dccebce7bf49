"""CPU tests of the host-side helpers either side of the hot path against golden
outputs of the reference's own functions (tests/golden/host_utils.json, oracle/make_golden.py
stage_hostutils)."""
import json
import os
import random
from collections import namedtuple

import numpy as np
import pytest

import gcrl_testutil as util
from graph_colouring_with_rl_b200 import utils as U
from graph_colouring_with_rl_b200 import dataset_functions as DF

with open(os.path.join(util.GOLD, 'host_utils.json')) as _f:
    GOLDEN = json.load(_f)


def _ind_triples(ind):
    return sorted([u, v, k] for u, dd in ind.items() for v, k in dd.items())


def test_dimacs_reader_matches_reference(tmp_path):
    fp = str(tmp_path / 'tiny.col')
    with open(fp, 'w') as f:
        f.write(GOLDEN['dimacs_text'])
    g = GOLDEN['dimacs']
    n, el, ea, ind, idx = U.import_dimacs_graph(fp)
    assert n == g['n'] and isinstance(el, tuple) and isinstance(ea, tuple) and isinstance(idx, tuple)
    assert [list(e) for e in el] == g['edge_list']
    assert list(ea) == g['edge_attr']
    assert _ind_triples(ind) == sorted(g['ind']) and set(ind.keys()) == set(range(n))
    assert sorted(idx) == g['indices']
    adj = U.import_dimacs_adjacency(fp)
    assert adj.dtype == np.uint8 and (adj == adj.T).all() and adj.diagonal().sum() == 0
    ela = np.asarray(g['edge_list'])
    assert (-(adj[ela[:, 0], ela[:, 1]].astype(int)) == np.asarray(g['edge_attr'])).all()
    nxg = U.import_dimacs_graph_as_networkx(fp)
    assert sorted(sorted(e) for e in nxg.edges()) == GOLDEN['dimacs_nx_edges']
    with open(fp, 'w') as f:
        f.write('c no problem line\ne 1 2\n')
    with pytest.raises(ValueError):
        U.import_dimacs_graph(fp)


def test_queen_graphs_match_reference_including_rng_draws():
    for q in GOLDEN['queens']:
        random.seed(q['seed'])
        el, ea, ind, idx = U.generate_queen_graph(q['n'], q['k'])
        assert [list(e) for e in el] == q['edge_list']
        assert list(ea) == q['edge_attr']
        assert list(idx) == q['indices']
        assert random.random() == q['after']            # exactly one draw consumed, like the reference
    with pytest.raises(Exception):
        U.generate_queen_graph(10, 3)


def test_leighton_graphs_match_reference_including_rng_draws():
    for g in GOLDEN['leighton']:
        random.seed(g['seed'])
        el, ea, ind, idx = U.generate_leighton_graph(g['n'], g['k'])
        assert list(ea) == g['edge_attr'] and list(idx) == g['indices']      # incl. the reference's set-iteration order
        assert random.random() == g['after']
        assert np.asarray(el).tolist() == U.complete_edge_arrays(g['n']).tolist()
        random.seed(g['seed'])
        adj = U.leighton_adjacency(g['n'], g['k'])
        ela = np.asarray(el)
        assert (-(adj[ela[:, 0], ela[:, 1]].astype(int)) == np.asarray(g['edge_attr'])).all()
    p = GOLDEN['leighton_params']
    random.seed(p['seed'])
    m, c, a, b = U.generate_leighton_parameters(p['n'], p['k'])
    assert (m, c, a, b) == (p['m'], p['c'], p['a'], p['b'])
    # Leighton's conditions: hcf(m, n) = k, hcf(c, m) = 1, every prime factor of m divides a - 1
    from math import gcd
    assert gcd(m, p['n']) == p['k'] and gcd(c, m) == 1
    assert all((a - 1) % q == 0 for q in set(U.get_prime_factors(m)))


def test_convert_networkx_matches_reference():
    c = GOLDEN['convert']
    el, ea, ind, idx = U.convert_networkx_to_graph(c['vertices'], [tuple(e) for e in c['edges']])
    assert [list(e) for e in el] == c['edge_list'] and isinstance(el, tuple)
    assert list(ea) == c['edge_attr'] and list(idx) == c['indices']
    assert _ind_triples(ind) == sorted(c['ind'])


def test_stat_files_byte_identical_and_round_trip(tmp_path):
    T = namedtuple("Stats", ["saved_episodes", "episode_rewards", "colours_used"])
    V = namedtuple("Stats", ["episode_rewards", "colours_used"])
    tr = T([1, 2, 3], [-7, -5, -6], [7, 5, 6])
    va = {0: V([-7.0, -5.0], [7.0, 5.0]), 500: V([-6.0, -4.0], [6.0, 4.0])}
    be = V([-8.25, -6.5], [8.25, 6.5])
    U.save_training_stats_to_file(tr, str(tmp_path / 't'))
    U.save_validation_stats_to_file(va, str(tmp_path / 'v'))
    U.save_benchmark_stats_to_file(be, str(tmp_path / 'b'))
    for k in 'tvb':
        assert open(str(tmp_path / k)).read() == GOLDEN['stat_files'][k]
    t2 = U.load_stats_from_file(str(tmp_path / 't'))
    assert t2.saved_episodes == [1.0, 2.0, 3.0] and t2.colours_used == [7.0, 5.0, 6.0]
    v2 = U.load_validation_stats_from_file(str(tmp_path / 'v'))
    assert sorted(v2) == [0, 500] and v2[500].episode_rewards == [-6.0, -4.0]
    b2 = U.load_benchmark_stats_from_file(str(tmp_path / 'b'))
    assert b2.episode_rewards == [-8.25, -6.5]


def test_edge_index_closed_form_and_encoding():
    for n in (2, 3, 7, 50):
        el = U.complete_edge_arrays(n)
        assert el.shape == (n * (n - 1), 2)
        assert (U.edge_index_of(el[:, 0], el[:, 1], n) == np.arange(el.shape[0])).all()
    g = util.load_validation_graphs()[5]
    el, ea, ind = U.encode_adjacency(g['adj'])
    ela = np.asarray(el)
    assert (ela == U.complete_edge_arrays(g['n'])).all()
    assert (-(g['adj'][ela[:, 0], ela[:, 1]].astype(int)) == np.asarray(ea)).all()
    assert ind[3][1] == int(U.edge_index_of(3, 1, g['n']))


def test_definitional_graphs():
    nx = pytest.importorskip('networkx')
    for k in range(2, 8):
        ref = nx.to_numpy_array(nx.mycielski_graph(k), dtype=np.uint8)
        assert (U.mycielski_adjacency(k) == ref).all()
    # DIMACS sizes: queen5_5 25 vertices / 160 edges (the file lists 320 directed), queen8_12 96 / 1368, myciel7 191 / 2360
    q = DF.definitional_graph('queen5_5')
    assert q.shape == (25, 25) and q.sum() // 2 == 160
    q = DF.definitional_graph('queen8_12')
    assert q.shape == (96, 96) and q.sum() // 2 == 1368
    m = DF.definitional_graph('myciel7')
    assert m.shape == (191, 191) and m.sum() // 2 == 2360
    q8 = nx.from_numpy_array(DF.definitional_graph('queen8_8'))
    assert q8.number_of_edges() == 728
    names = [nm for nm, _ in DF.lemos_reconstructible()]
    assert len(names) == 13 and 'queen13_13' in names
    with pytest.raises(KeyError):
        DF.definitional_graph('games120')
    rng = np.random.default_rng(0)
    a = U.erdos_renyi_adjacency(40, 0.5, rng)
    assert (a == a.T).all() and a.diagonal().sum() == 0 and 250 < a.sum() // 2 < 530


def test_dataset_unpickler_handles_main_dataset(tmp_path):
    import pickle
    import sys
    import types
    # a pickle that names __main__.Dataset, like the reference's dataset files
    main = sys.modules['__main__']
    had = hasattr(main, 'Dataset')
    old = getattr(main, 'Dataset', None)
    cls = type('Dataset', (), {})
    cls.__module__ = '__main__'
    main.Dataset = cls
    try:
        obj = cls()
        obj.graphs = [{'name': 'g0', 'no_vertices': 3}]
        obj.dsatur_stats_combined = {'avg_colours_used': 2.0}
        fp = str(tmp_path / 'ds.pickle')
        with open(fp, 'wb') as f:
            pickle.dump(obj, f)
    finally:
        if had:
            main.Dataset = old
        else:
            del main.Dataset
    ds = DF.load_dataset(fp)
    assert isinstance(ds, DF.Dataset) and ds.graphs[0]['name'] == 'g0' and ds.dsatur_stats_combined['avg_colours_used'] == 2.0
