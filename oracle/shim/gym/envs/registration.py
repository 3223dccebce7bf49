import importlib

_REG = {}


def register(id, entry_point, **kw):
    _REG[id] = entry_point


def make(id, **kwargs):
    # the rollout harness asks for 'graph-colouring' without the version suffix
    # (test_policy_functions.py:13); resolve by prefix.
    key = id if id in _REG else next(k for k in _REG if k.startswith(id))
    mod, cls = _REG[key].split(':')
    return getattr(importlib.import_module(mod), cls)(**kwargs)
