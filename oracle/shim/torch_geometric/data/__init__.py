import torch


class Data(object):
    """x [N,Fv], edge_index [2,E] int64, edge_attr [E,Fe]; `.to()`/`.cpu()` move all three."""

    def __init__(self, x=None, edge_index=None, edge_attr=None):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr

    def to(self, device):
        return Data(self.x.to(device), self.edge_index.to(device), self.edge_attr.to(device))

    def cpu(self):
        return self.to('cpu')


class Batch(Data):
    @staticmethod
    def from_data_list(data_list):
        """PyG collate: cat x / edge_attr on dim 0, edge_index on dim 1 offset by the
        cumulative node count; `.batch` = vertex -> graph id."""
        xs, eis, eas, bs = [], [], [], []
        off = 0
        for g, d in enumerate(data_list):
            xs.append(d.x)
            eis.append(d.edge_index + off)
            eas.append(d.edge_attr)
            bs.append(torch.full((d.x.shape[0],), g, dtype=torch.long))
            off += d.x.shape[0]
        b = Batch(torch.cat(xs, 0), torch.cat(eis, 1), torch.cat(eas, 0))
        b.batch = torch.cat(bs, 0)
        b.num_graphs = len(data_list)
        return b

    def to(self, device):
        b = Batch(self.x.to(device), self.edge_index.to(device), self.edge_attr.to(device))
        if hasattr(self, 'batch'):
            b.batch = self.batch.to(device)
            b.num_graphs = self.num_graphs
        return b
