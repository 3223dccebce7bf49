import inspect
import torch


class _Inspector(object):
    def __init__(self, module):
        self.module = module

    def distribute(self, func_name, coll_dict):
        """PyG <= 2.2 `Inspector.distribute`: bind every named parameter of `message`, and
        every parameter but the first positional of `aggregate` / `update`, from the
        collected dict, falling back to the parameter default (this is how `dim_dim`
        becomes None at architecture.py:62,89-91)."""
        params = list(inspect.signature(getattr(self.module, func_name)).parameters.values())
        if func_name != 'message':
            params = params[1:]
        out = {}
        for p in params:
            if p.name in coll_dict:
                out[p.name] = coll_dict[p.name]
            elif p.default is not inspect.Parameter.empty:
                out[p.name] = p.default
            else:
                raise TypeError('Required parameter %s is empty.' % p.name)
        return out


class MessagePassing(torch.nn.Module):
    """flow = source_to_target: x_j = x[edge_index[0]], x_i = x[edge_index[1]],
    index = edge_index[1] (aggregation at the target)."""

    def __init__(self, aggr=None):
        super().__init__()
        self.aggr = aggr
        self.inspector = _Inspector(self)
        self.__user_args__ = None

    def __check_input__(self, edge_index, size):
        return [None, None] if size is None else list(size)

    def __collect__(self, args, edge_index, size, kwargs):
        out = dict(kwargs)
        x = kwargs.get('x')
        if x is not None:
            out['x_j'] = x.index_select(0, edge_index[0])
            out['x_i'] = x.index_select(0, edge_index[1])
        out['edge_index'] = edge_index
        out['index'] = edge_index[1]
        out['size'] = size
        return out
