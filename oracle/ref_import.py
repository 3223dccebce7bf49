"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference (/root/reference) in this
container, with the third-party packages it needs (torch_scatter, torch_geometric, gym --
none installed, no network) replaced by the stand-ins in oracle/shim/.

/root/reference does not exist on the GPU box, so nothing here is used by `-m gpu` tests or
smoke(); it exists to (1) pin oracle/gn_oracle.py + oracle/env_oracle.py against the real
reference, (2) generate the committed fixtures under tests/golden/ (oracle/make_golden.py) and
(3) let bench.py's CPU baseline legs (`--impl reference`, `cpu_baseline`) time the reference's
own modules from the byte-for-byte copy oracle/build_ref.py leaves in oracle/_ref/.
"""
import os
import pickle
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get('GCRL_REFERENCE_ROOT', '/root/reference')
if not os.path.isfile(os.path.join(REFERENCE_ROOT, 'architecture.py')):
    # the GPU box: the reference's own files as copied by oracle/build_ref.py (git-ignored, travels with gpurun)
    REFERENCE_ROOT = os.path.join(_HERE, '_ref')
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shim')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'architecture.py'))


def import_reference():
    """Returns a namespace with the reference's own modules (architecture, pna, dqn_agent,
    env, utils, test_policy_functions)."""
    if not reference_available():
        raise RuntimeError('reference checkout not present at %s' % REFERENCE_ROOT)
    os.environ.setdefault('CUDA_VISIBLE_DEVICES', '')  # architecture.py:11 pins cuda:4
    for p in (_SHIM, REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, 'gym-graph-colouring')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import types
    import architecture
    import principal_neighbourhood_aggregation as pna
    import dqn_agent
    import utils as ref_utils
    import gym_graph_colouring  # registers graph-colouring-v0
    from gym_graph_colouring.envs import graph_colouring_env
    import test_policy_functions
    ns = types.SimpleNamespace(architecture=architecture, pna=pna, dqn_agent=dqn_agent,
                               utils=ref_utils, env=graph_colouring_env,
                               test_policy_functions=test_policy_functions)
    return ns


class DatasetHolder(object):
    """The pickles reference `__main__.Dataset`, whose real constructor is broken
    (dataset_functions.py:42-57); unpickle into this plain holder instead."""
    pass


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if name == 'Dataset':
            return DatasetHolder
        return super().find_class(module, name)


def load_reference_dataset(name='validation_dataset.pickle'):
    with open(os.path.join(REFERENCE_ROOT, 'datasets', name), 'rb') as f:
        return _Unpickler(f).load()


def reference_target(ref, adj, name='graph'):
    """Target dict of a 0/1 adjacency matrix assembled with the reference's OWN helpers (utils.py:296-323
    convert_networkx_to_graph, :515-565 distances / counts): the keys the environment and the rollout harness read
    (SURVEY 8f N2), without construct_target's 40 networkx centralities (seconds per graph, never read by the path)."""
    import numpy as np
    U = ref.utils
    n = int(adj.shape[0])
    iu, ju = np.nonzero(np.triu(adj, 1))
    edges = [(int(a), int(b)) for a, b in zip(iu, ju)]
    edge_list, edge_attr, edge_ind_dict, target_edge_indices = U.convert_networkx_to_graph(range(n), edges)
    present_edges = tuple(edge_list[i] for i in target_edge_indices)
    adjm = U.get_adj_matrix(n, present_edges)
    dist = U.get_neighbour_distances(n, present_edges, adjm)
    return {'name': name, 'no_vertices': n, 'vertex_names': list(range(n)), 'edge_list': edge_list,
            'edge_attr': edge_attr, 'edge_ind_dict': edge_ind_dict, 'target_edge_indices': target_edge_indices,
            'present_edges': present_edges, 'neighbour_distances': dist,
            'neighbour_dist_counts': U.get_neighbour_distance_counts(n, dist)}
