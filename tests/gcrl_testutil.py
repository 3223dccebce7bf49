"""Shared fixtures for the tests: the committed golden data (tests/golden/, produced by
oracle/make_golden.py from the unmodified reference) and small graph builders."""
import os
from collections import OrderedDict

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_validation_graphs():
    """-> list of dict(n, adj [n,n] uint8, edge_list [E,2] int64, name, dsatur_order, dsatur_colours,
    dist_counts [n,3])."""
    d = np.load(os.path.join(GOLD, 'validation_graphs.npz'))
    out, ao, eo, oo, co = [], 0, 0, 0, 0
    for i, n in enumerate(d['n']):
        n = int(n)
        E = n * (n - 1)
        adj = d['adj'][ao:ao + n * n].reshape(n, n).astype(np.uint8)
        el = d['edge_list'][eo:eo + 2 * E].reshape(E, 2).astype(np.int64)
        ol = int(d['dsatur_order_len'][i])
        out.append(dict(n=n, adj=adj, edge_list=el, name=str(d['names'][i]),
                        dsatur_order=d['dsatur_order'][oo:oo + ol].astype(np.int64),
                        dsatur_colours=int(d['dsatur_colours_used'][i]),
                        dist_counts=d['neighbour_dist_counts'][co:co + 3 * n].reshape(n, 3).astype(np.int64)))
        ao += n * n
        eo += 2 * E
        oo += ol
        co += 3 * n
    return out


def load_parity_weights():
    sd = torch.load(os.path.join(GOLD, 'parity_weights_GN'), map_location='cpu')
    return OrderedDict((k, v.float()) for k, v in sd.items())


def flatten(params):
    return torch.cat([v.reshape(-1) for v in params.values()])


def unflatten(flat, like):
    out, o = OrderedDict(), 0
    for k, v in like.items():
        out[k] = flat[o:o + v.numel()].view(v.shape).clone()
        o += v.numel()
    return out


def split_by(arr, lens):
    out, o = [], 0
    for l in lens:
        out.append(arr[o:o + int(l)])
        o += int(l)
    return out


def random_colours(n, adj, frac, rng):
    """A valid partial first-fit colouring of about frac*n vertices (-1 = uncoloured)."""
    col = -np.ones(n, dtype=np.int64)
    for v in rng.permutation(n)[:max(0, int(frac * n))]:
        used = set(int(c) for c in col[adj[v].astype(bool)] if c >= 0)
        c = 0
        while c in used:
            c += 1
        col[v] = c
    return col
