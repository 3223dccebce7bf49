"""Drop-in for the parts of the reference's dataset_functions.py the hot path's callers use
(:125-149): `run_learned_policy_on_dataset`, `load_dataset`, plus dataset assembly for the
graph families that can be rebuilt by definition (the Lemos pickle and the DIMACS
files are absent from the checkout; queen and Mycielski graphs are definitional).
"""
import pickle

import numpy as np

from .utils import mycielski_adjacency, queen_adjacency


class Dataset(object):
    """Plain holder: `.graphs` plus whatever stats the pickle carries.  (The reference class's
    constructor calls functions that no longer exist, dataset_functions.py:42-57; its pickles
    name the class `__main__.Dataset`.)"""

    def __init__(self, graphs=None):
        self.graphs = graphs if graphs is not None else []


class _DatasetUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if name == 'Dataset':
            return Dataset
        return super().find_class(module, name)


def load_dataset(filepath):
    """dataset_functions.py:147-149, robust to the `__main__.Dataset` class path."""
    with open(filepath, 'rb') as f:
        return _DatasetUnpickler(f).load()


def run_learned_policy_on_dataset(graphs, agent, stochastic=False):
    """dataset_functions.py:125-136: one batched device rollout over all graphs."""
    from .test_policy_functions import test_using_learned_policy
    return test_using_learned_policy(graphs, agent, stochastic=stochastic)


# DIMACS names -> (rows, cols) of the queen graphs in the Lemos set
QUEEN_BOARDS = {'queen5_5': (5, 5), 'queen6_6': (6, 6), 'queen7_7': (7, 7), 'queen8_8': (8, 8),
                'queen8_12': (8, 12), 'queen9_9': (9, 9), 'queen10_10': (10, 10), 'queen11_11': (11, 11),
                'queen12_12': (12, 12), 'queen13_13': (13, 13)}
MYCIEL_ORDER = {'myciel3': 4, 'myciel4': 5, 'myciel5': 6, 'myciel6': 7, 'myciel7': 8}


def definitional_graph(name):
    """Adjacency of a DIMACS benchmark graph that is defined by construction."""
    if name in QUEEN_BOARDS:
        return queen_adjacency(*QUEEN_BOARDS[name])
    if name in MYCIEL_ORDER:
        return mycielski_adjacency(MYCIEL_ORDER[name])
    raise KeyError('%s cannot be rebuilt offline (needs its DIMACS file)' % name)


def lemos_reconstructible(names=None, as_targets=False, device=None):
    """The subset of the Lemos 20-graph benchmark that can be rebuilt without files, as
    (name, adjacency) pairs or hot-path target dicts (envs.make_target)."""
    names = list(QUEEN_BOARDS) + ['myciel5', 'myciel6', 'myciel7'] if names is None else names
    out = []
    for nm in names:
        adj = definitional_graph(nm)
        if as_targets:
            from .envs import make_target
            out.append(make_target(adj, name=nm, device=device))
        else:
            out.append((nm, np.ascontiguousarray(adj)))
    return out
