"""Drop-in for the reference's architecture.py: `GNNBlock` and `GNNetwork` with the same
constructor arguments, attributes (`name`, `optimizer`, `device`, `gn_layer1..5`, `fc1..3`),
`state_dict` key schema (46 tensors) and `forward` return tuple, executed by the hand-written
sm_100a kernels of libgcrl_b200 through the C ABI (include/gcrl_b200.h).

Reference: architecture.py:13-99 (GNNBlock), :102-213 (GNNetwork).

How the module maps onto the library: all 46 parameters are views into ONE flat fp32 buffer in
state_dict order (the layout `gcrl_param_layout` describes), so the kernels read the weights
in place and `optimizer.step()` is one fused launch over 198 529 floats.  `forward` is
differentiable w.r.t. the parameters through a custom autograd function that calls
`gcrl_gn_forward` (with activation stash) / `gcrl_gn_backward`.
"""
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib, ops
from .data import Data
from .principal_neighbourhood_aggregation import AGGREGATORS, SCALERS

# The reference pins "cuda:4" (architecture.py:11).  Here: one process per GPU, the process's
# device (LOCAL_RANK under torchrun), overridable with GCRL_DEVICE.
DEVICE = os.environ.get('GCRL_DEVICE', 'cuda:%d' % int(os.environ.get('LOCAL_RANK', '0')))

MATH_MODES = {'fp32': _lib.GCRL_MATH_FP32, 'bf16': _lib.GCRL_MATH_BF16, 'bf16x3': _lib.GCRL_MATH_BF16X3}


def _device():
    if not torch.cuda.is_available():
        raise _lib.GcrlError('graph_colouring_with_rl_b200 needs a CUDA device (sm_100a); there is no CPU path')
    return torch.device(DEVICE)


class _PackCache(object):
    """Segment layouts keyed by the identity of the caller's edge_index tensor: the environment
    hands out the same static edge_index every step of an episode (graph_colouring_env.py:136),
    so batching is done once per graph, not once per step.

    Bounded by BYTES (layout + permuted edge_attr + the key tensors it keeps alive), default 256 MB, and
    single inputs larger than a quarter of the budget are not cached at all: a caller that builds a fresh
    `Batch.from_data_list(...)` per learn() never hits the cache and must not pin a minibatch per entry.
    The key includes the tensors' `_version`; writes that bypass it (`.data`, raw-pointer kernels) are not
    seen -- do not mutate a cached edge_index / edge_attr in place, or call `clear_pack_cache()`."""

    def __init__(self, capacity=256, max_bytes=256 << 20):
        self.capacity = capacity
        self.max_bytes = int(max_bytes)
        self.bytes = 0
        self.items = OrderedDict()

    @staticmethod
    def _nbytes(*tensors):
        return sum(t.numel() * t.element_size() for t in tensors if t is not None)

    def clear(self):
        self.items.clear()
        self.bytes = 0

    def get(self, edge_index, n_vertices, edge_attr):
        key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(n_vertices),
               edge_attr.data_ptr(), edge_attr._version)
        hit = self.items.get(key)
        if hit is not None:
            self.items.move_to_end(key)
            return hit[0], hit[1]
        packed = ops.pack(edge_index, n_vertices)
        eattr_i = ops.permute_rows(edge_attr.reshape(edge_attr.shape[0], -1), packed.perm, True)
        nb = self._nbytes(packed.seg_off, packed.src, packed.dst, packed.perm, eattr_i, edge_index, edge_attr)
        if nb > self.max_bytes // 4:
            return packed, eattr_i
        # keep the key tensors alive so their addresses cannot be recycled while cached
        self.items[key] = (packed, eattr_i, edge_index, edge_attr, nb)
        self.bytes += nb
        while len(self.items) > self.capacity or self.bytes > self.max_bytes:
            _, old = self.items.popitem(last=False)
            self.bytes -= old[4]
        return packed, eattr_i


_pack_cache = _PackCache()


def clear_pack_cache():
    """Drop every cached segment layout (after mutating an edge_index / edge_attr in place, or to free memory)."""
    _pack_cache.clear()


def _mlp(in_dim, hidden_dim, out_dim):
    return nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, out_dim))


class GNNBlock(nn.Module):
    """One message / aggregate / update block (architecture.py:13-99).

    `forward(data) -> (msg_out [E,64] in the caller's edge order, upd_out [N,64])`; the
    aggregation is the fixed PNA set mean|min|max|std with identity scaling, which is what the
    reference's `aggregate` evaluates (its scalers are stored and never applied, :31,:62-66).
    """

    def __init__(self, message_in_dim, message_hidden_dim, message_out_dim, update_in_dim, update_hidden_dim,
                 update_out_dim, aggregators, scalers):
        super().__init__()
        if (message_hidden_dim, message_out_dim, update_hidden_dim, update_out_dim) != (64, 64, 64, 64):
            raise ValueError('libgcrl_b200 implements the reference widths only (hidden = out = 64)')
        if list(aggregators) != ['mean', 'min', 'max', 'std']:
            raise ValueError("libgcrl_b200 fuses the reference's aggregator set ['mean','min','max','std']")
        self.aggregators = [AGGREGATORS[a] for a in aggregators]
        self.scalers = [SCALERS[s] for s in scalers]
        self.vertex_dim = update_in_dim - 4 * message_out_dim
        self.edge_dim = message_in_dim - 2 * self.vertex_dim
        if self.vertex_dim < 1 or self.edge_dim < 1:
            raise ValueError('inconsistent message/update input widths')
        self.message_mlp = _mlp(message_in_dim, message_hidden_dim, message_out_dim)
        self.update_mlp = _mlp(update_in_dim, update_hidden_dim, update_out_dim)
        self.device = _device()
        self.to(self.device)

    def forward(self, data):
        """Stand-alone block evaluation (inference; training goes through GNNetwork.forward)."""
        x = data.x.to(self.device, torch.float32)
        ei = data.edge_index.to(self.device)
        ea = data.edge_attr.to(self.device, torch.float32)
        packed, ea_i = _pack_cache.get(ei, x.shape[0], ea)
        flat = torch.cat([p.detach().reshape(-1) for p in self.parameters()])
        # a network whose first block is this block: same flat prefix
        msg_i, upd = ops.gn_block_forward(flat, (self.vertex_dim, self.edge_dim, 0), 1, packed, x, ea_i)
        return ops.permute_rows(msg_i, packed.perm, False), upd


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8) semantics (architecture.py:173) as ONE
    launch of gcrl_adam_soft_update over the network's flat parameter buffer; when a target
    buffer is attached the DQN soft update (dqn_agent.py:226-229) rides in the same launch."""

    def __init__(self, net, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(list(net.parameters()), dict(lr=lr, betas=betas, eps=eps))
        self._net = net
        self.step_count = 0
        self.exp_avg = None
        self.exp_avg_sq = None

    def _flat_grad(self):
        net = self._net
        flat = net.flat_params()
        g = net._flat_grad
        if g is None or g.device != flat.device:
            g = net._flat_grad = torch.zeros_like(flat)
        base = g.data_ptr()
        for p, off in zip(net._param_list(), net._offsets):
            if p.grad is None:
                g[off:off + p.numel()].zero_()
            elif p.grad.data_ptr() != base + 4 * off:
                g[off:off + p.numel()].copy_(p.grad.reshape(-1))
        return g

    @torch.no_grad()
    def step(self, closure=None, target_flat=None, tau=0.0, grad=None):
        loss = closure() if closure is not None else None
        flat = self._net.flat_params(full_check=grad is None)
        if grad is None:
            grad = self._flat_grad()
        if self.exp_avg is None or self.exp_avg.device != flat.device:
            self.exp_avg = torch.zeros_like(flat)
            self.exp_avg_sq = torch.zeros_like(flat)
        self.step_count += 1
        grp = self.param_groups[0]
        ops.adam_soft_update(flat, grad, self.exp_avg, self.exp_avg_sq, target_flat, self.step_count,
                             lr=grp['lr'], betas=grp['betas'], eps=grp['eps'], tau=tau)
        return loss

    def state_dict(self):
        return dict(step=self.step_count, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq,
                    param_groups=[{k: v for k, v in g.items() if k != 'params'} for g in self.param_groups])

    def load_state_dict(self, sd):
        self.step_count = int(sd['step'])
        dev = self._net.flat_params().device
        self.exp_avg = None if sd['exp_avg'] is None else sd['exp_avg'].to(dev).clone()
        self.exp_avg_sq = None if sd['exp_avg_sq'] is None else sd['exp_avg_sq'].to(dev).clone()
        for g, s in zip(self.param_groups, sd.get('param_groups', [])):
            g.update(s)


class _GNFunction(torch.autograd.Function):
    """(h, q) = GN(params; graph batch) with the hand-written backward."""

    @staticmethod
    def forward(ctx, net, packed, x, eattr_i, gfeat, vgraph, *params):
        flat = net.flat_params()
        h, q, stash = ops.gn_forward(flat, net.dims, packed, x, eattr_i, gfeat, vgraph, with_stash=True,
                                     math_mode=net.math_mode)
        ctx.net, ctx.packed, ctx.stash = net, packed, stash
        ctx.gfeat, ctx.vgraph = gfeat, vgraph
        ctx.save_for_backward(x, eattr_i, flat)
        return h, q.view(-1, 1)

    @staticmethod
    def backward(ctx, d_h, d_q):
        x, eattr_i, flat = ctx.saved_tensors
        net = ctx.net
        if d_h is None and d_q is None:
            return (None,) * (6 + len(net._offsets))
        grad = ops.gn_backward(flat, net.dims, ctx.packed, x, eattr_i, ctx.stash,
                               d_q=None if d_q is None else d_q.reshape(-1).float(),
                               d_h=None if d_h is None else d_h.float(), gfeat=ctx.gfeat, vgraph=ctx.vgraph,
                               math_mode=net.math_mode)
        views = tuple(grad[off:off + p.numel()].view(p.shape) for p, off in zip(net._param_list(), net._offsets))
        return (None, None, None, None, None, None) + views


class GNNetwork(nn.Module):
    """Five GNNBlocks followed by fc1/fc2/fc3 (architecture.py:102-213)."""

    def __init__(self, vertex_features_dim, edge_features_dim, global_features_dim, name, alpha, math_mode='fp32'):
        super().__init__()
        self.name = name
        aggregators = ['mean', 'min', 'max', 'std']      # architecture.py:119
        scalers = ['identity']                           # architecture.py:120
        hid = 64                                         # architecture.py:123-128
        n_agg = len(aggregators) * len(scalers)
        self.dims = (int(vertex_features_dim), int(edge_features_dim), int(global_features_dim))
        self.math_mode = MATH_MODES[math_mode] if isinstance(math_mode, str) else int(math_mode)
        dv, de = self.dims[0], self.dims[1]
        for l in range(1, 6):                            # creation order fixes the RNG draws (:157-161)
            setattr(self, 'gn_layer%d' % l, GNNBlock(2 * dv + de, hid, hid, n_agg * hid + dv, hid, hid,
                                                     aggregators, scalers))
            dv, de = hid, hid
        self.fc1 = nn.Linear(hid + self.dims[2], hid)    # :165-167
        self.fc2 = nn.Linear(hid, hid)
        self.fc3 = nn.Linear(hid, 1)                     # actions_per_vertex = 1 (:129)
        self.dummy_data = Data(x=torch.tensor([]), edge_index=torch.tensor([], dtype=torch.long),
                               edge_attr=torch.tensor([]))
        self.device = _device()
        self.to(self.device)
        self._offsets = ops.param_layout(*self.dims)[:-1]
        self._n_params = ops.param_layout(*self.dims)[-1]
        self._flat = None
        self._flat_grad = None
        self.flat_params()
        self.optimizer = FusedAdam(self, lr=alpha)       # :173

    # ---- flat parameter buffer ------------------------------------------------------------
    def _param_list(self):
        """The 46 parameters in state_dict order, cached: nn.Module.parameters() walks the module tree (~0.1 ms), and the
        hot loops ask for the flat buffer several times per step.  The Parameter OBJECTS never change (load_state_dict,
        .to() and optimizers rewrite their .data), so the list stays valid; flat_params() re-checks the storages."""
        pl = self.__dict__.get('_params_cached')
        if pl is None:
            pl = list(self.parameters())
            self.__dict__['_params_cached'] = pl
        return pl

    def flat_params(self, full_check=True):
        """The [198 529] fp32 buffer all parameters alias.  Rebuilt (and the parameters
        re-pointed) if anything replaced a parameter's storage, e.g. `.to()` or `.data = ...`.
        full_check=False (the learner's and the rollouts' inner loops, which own the network between calls) looks at the
        first and the last parameter only: module-wide moves (`.to()`, `load_state_dict(assign=True)`) are still seen."""
        params = self._param_list()
        flat = self._flat
        if flat is not None:
            base = flat.data_ptr()
            if not full_check:
                if params[0].data_ptr() == base and params[-1].data_ptr() == base + 4 * self._offsets[-1]:
                    return flat
            elif all(p.data_ptr() == base + 4 * off for p, off in zip(params, self._offsets)):
                return flat
        assert [p.numel() for p in params] == [b - a for a, b in zip(self._offsets, self._offsets[1:] + [self._n_params])], \
            'parameter order does not match the library layout'
        dev = params[0].device
        if dev.type != 'cuda':
            raise _lib.GcrlError('GNNetwork parameters must live on a CUDA device')
        flat = torch.empty(self._n_params, dtype=torch.float32, device=dev)
        for p, off in zip(params, self._offsets):
            view = flat[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self._flat = flat
        self._flat_grad = None
        return flat

    # ---- forward ----------------------------------------------------------------------------
    def _prepare(self, data, global_features, vertex_batch_map):
        x = data.x
        _lib.require_cuda(x, 'data.x')
        x = x.float()
        packed, eattr_i = _pack_cache.get(data.edge_index, x.shape[0], data.edge_attr.float())
        gfeat = vgraph = None
        if self.dims[2] > 0:
            gfeat = torch.as_tensor(global_features, dtype=torch.float32, device=x.device).reshape(-1, self.dims[2]).contiguous()
            vgraph = torch.as_tensor(vertex_batch_map, dtype=torch.int32, device=x.device).contiguous()
        return x, packed, eattr_i, gfeat, vgraph

    def forward(self, data, global_features, vertex_batch_map, edge_batch_map, called_by=None):
        """-> (vertex_embeddings [N,64], q_values [N,1]) (architecture.py:179-199).
        `edge_batch_map` and `called_by` are accepted and unused, as in the reference."""
        x, packed, eattr_i, gfeat, vgraph = self._prepare(data, global_features, vertex_batch_map)
        params = self._param_list()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _GNFunction.apply(self, packed, x, eattr_i, gfeat, vgraph, *params)
        h, q, _ = ops.gn_forward(self.flat_params(), self.dims, packed, x, eattr_i, gfeat, vgraph,
                                 math_mode=self.math_mode)
        return h, q.view(-1, 1)

    # ---- checkpoints (architecture.py:201-213) ------------------------------------------------
    def save_checkpoint(self, checkpoint_filepath_root):
        print('... saving checkpoint ...')
        checkpoint_filepath_full = checkpoint_filepath_root + '_' + self.name
        torch.save(OrderedDict((k, v.detach().cpu().clone()) for k, v in self.state_dict().items()),
                   checkpoint_filepath_full)
        return checkpoint_filepath_full

    def load_checkpoint(self, checkpoint_filepath_root):
        if checkpoint_filepath_root.endswith(self.name):
            checkpoint_filepath_full = checkpoint_filepath_root
        else:
            checkpoint_filepath_full = checkpoint_filepath_root + '_' + self.name
        print('... loading checkpoint from ' + checkpoint_filepath_full + ' ...')
        self.load_state_dict(torch.load(checkpoint_filepath_full, map_location=self.device))
        self.flat_params()
