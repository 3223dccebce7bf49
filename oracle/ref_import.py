"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference (/root/reference) in this
container, with the third-party packages it needs (torch_scatter, torch_geometric, gym --
none installed, no network) replaced by the stand-ins in oracle/shim/.

/root/reference does not exist on the GPU box, so nothing here is used by `-m gpu` tests,
smoke() or bench.py; it exists to (1) pin oracle/gn_oracle.py + oracle/env_oracle.py against
the real reference and (2) generate the committed fixtures under tests/golden/
(oracle/make_golden.py).
"""
import os
import pickle
import sys

REFERENCE_ROOT = os.environ.get('GCRL_REFERENCE_ROOT', '/root/reference')
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
