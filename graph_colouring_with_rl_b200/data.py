"""Minimal graph containers with the attribute surface of torch_geometric.data.Data / Batch that
the reference touches (architecture.py:7,182-187; dqn_agent.py:5,55-56,122,160,163;
graph_colouring_env.py:139).  torch_geometric is not a dependency; objects from a real PyG
install are accepted everywhere by duck typing (x / edge_index / edge_attr / to / cpu)."""
import torch


class Data(object):
    def __init__(self, x=None, edge_index=None, edge_attr=None, **extra):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        for k, v in extra.items():
            setattr(self, k, v)

    def _tensor_items(self):
        return [(k, v) for k, v in self.__dict__.items() if torch.is_tensor(v)]

    def to(self, device, non_blocking=False):
        """Moves every tensor attribute IN PLACE and returns self (PyG semantics)."""
        for k, v in self._tensor_items():
            setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self

    def cpu(self):
        return self.to('cpu')

    def cuda(self, device=None):
        return self.to('cuda' if device is None else device)

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        out.__dict__.update({k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.__dict__.items()})
        return out

    @property
    def num_nodes(self):
        return int(self.x.shape[0])

    @property
    def num_edges(self):
        return int(self.edge_index.shape[1])

    def __repr__(self):
        parts = ['%s=%s' % (k, list(v.shape)) for k, v in self._tensor_items()]
        return '%s(%s)' % (self.__class__.__name__, ', '.join(parts))


class Batch(Data):
    """Block-diagonal union of graphs: x / edge_attr concatenated, edge_index shifted by the
    cumulative vertex count; `batch` maps vertices to graphs, `ptr` holds vertex offsets."""

    @classmethod
    def from_data_list(cls, data_list):
        xs, eis, eas, batch, ptr, off = [], [], [], [], [0], 0
        for g, d in enumerate(data_list):
            n = int(d.x.shape[0])
            xs.append(d.x)
            eis.append(d.edge_index + off)
            if d.edge_attr is not None:
                eas.append(d.edge_attr)
            batch.append(torch.full((n,), g, dtype=torch.long, device=d.x.device))
            off += n
            ptr.append(off)
        out = cls(x=torch.cat(xs, 0), edge_index=torch.cat(eis, 1),
                  edge_attr=torch.cat(eas, 0) if eas else None)
        out.batch = torch.cat(batch, 0)
        out.ptr = torch.tensor(ptr, dtype=torch.long, device=out.x.device)
        out.num_graphs = len(data_list)
        return out
