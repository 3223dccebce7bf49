"""Tensor-level wrappers over the C ABI (include/gcrl_b200.h).  Every function takes CUDA
tensors, enqueues on the current torch stream and returns torch tensors; nothing here has a
CPU implementation."""
from collections import namedtuple

import torch

from . import _lib
from ._lib import check, device_guard, ptr, stream_ptr, workspace, require_cuda

HID = 64
N_BLOCKS = 5

Packed = namedtuple('Packed', 'seg_off src dst perm n_vertices n_edges')


def sm_count():
    return _lib.load().gcrl_sm_count()


def param_layout(vertex_dim=2, edge_dim=1, global_dim=0):
    """Offsets (47 ints) of the 46 state_dict tensors inside the flat parameter buffer."""
    import ctypes
    off = (ctypes.c_int64 * 47)()
    check(_lib.load().gcrl_param_layout(vertex_dim, edge_dim, global_dim, off), 'gcrl_param_layout')
    return list(off)


def _i32(n, device):
    return torch.empty(int(n), dtype=torch.int32, device=device)


# ---- graph batching ---------------------------------------------------------------------
def pack(edge_index, n_vertices, validate=False):
    """edge_index [2,E] int64 (CUDA, row 0 = source, row 1 = target) -> Packed."""
    lib = _lib.load()
    require_cuda(edge_index, 'edge_index')
    if edge_index.dtype != torch.int64:
        raise TypeError('edge_index must be int64')
    ei = edge_index.contiguous()
    dev = ei.device
    E, N = int(ei.shape[1]), int(n_vertices)
    seg_off, src, dst, perm = _i32(N + 1, dev), _i32(E, dev), _i32(E, dev), _i32(E, dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev) if validate else None
    nb = lib.gcrl_pack_workspace_bytes(N, E)
    ws = workspace(nb, dev, 'pack')
    with device_guard(dev):
        check(lib.gcrl_pack(ptr(ei), E, N, ptr(seg_off), ptr(src), ptr(dst), ptr(perm), ptr(err), ptr(ws),
                            ws.numel(), stream_ptr(dev)), 'gcrl_pack')
    if validate and int(err.item()) != 0:
        raise IndexError('edge_index holds a vertex id outside [0, %d)' % N)
    return Packed(seg_off, src, dst, perm, N, E)


def pack_dense(adj, adj_off, node_off, n_vertices, n_edges):
    """Complete-graph batch from dense byte adjacency -> (Packed, eattr [E,1] internal order)."""
    lib = _lib.load()
    require_cuda(adj, 'adj')
    dev = adj.device
    G = int(node_off.numel()) - 1
    N, E = int(n_vertices), int(n_edges)
    seg_off, src, dst, perm = _i32(N + 1, dev), _i32(E, dev), _i32(E, dev), _i32(E, dev)
    eattr = torch.empty(E, 1, dtype=torch.float32, device=dev)
    ws = workspace(lib.gcrl_pack_dense_workspace_bytes(G), dev, 'pack')
    with device_guard(dev):
        check(lib.gcrl_pack_dense(ptr(adj), ptr(adj_off), ptr(node_off), G, N, E, ptr(seg_off), ptr(src),
                                  ptr(dst), ptr(perm), ptr(eattr), ptr(ws), ws.numel(), stream_ptr(dev)),
              'gcrl_pack_dense')
    return Packed(seg_off, src, dst, perm, N, E), eattr


def permute_rows(t, perm, to_internal):
    """to_internal: out[s] = t[perm[s]]; else out[perm[s]] = t[s]."""
    lib = _lib.load()
    require_cuda(t)
    t = t.contiguous().float()
    rows = int(t.shape[0])
    width = int(t.numel() // max(rows, 1)) if rows else 1
    out = torch.empty_like(t)
    with device_guard(t.device):
        check(lib.gcrl_permute_rows(ptr(t), ptr(out), ptr(perm), rows, max(width, 1), 1 if to_internal else 0,
                                    stream_ptr(t.device)), 'gcrl_permute_rows')
    return out


# ---- PNA aggregation ----------------------------------------------------------------------
def pna_forward(msg, seg_off, n_vertices, want_backward_state=False):
    """msg [E,F] internal order -> out [N,4F] (+ arg_min, arg_max, var when asked)."""
    lib = _lib.load()
    require_cuda(msg, 'msg')
    msg = msg.contiguous()
    dev = msg.device
    E, F = int(msg.shape[0]), int(msg.shape[1])
    N = int(n_vertices)
    out = torch.empty(N, 4 * F, dtype=torch.float32, device=dev)
    amin = amax = var = None
    if want_backward_state:
        amin = torch.empty(N, F, dtype=torch.int32, device=dev)
        amax = torch.empty(N, F, dtype=torch.int32, device=dev)
        var = torch.empty(N, F, dtype=torch.float32, device=dev)
    with device_guard(dev):
        check(lib.gcrl_pna_forward(ptr(msg), ptr(seg_off), N, E, F, ptr(out), ptr(amin), ptr(amax), ptr(var),
                                   stream_ptr(dev)), 'gcrl_pna_forward')
    if want_backward_state:
        return out, amin, amax, var
    return out


def pna_backward(msg, out, var, d_out, seg_off, dst, amin, amax):
    lib = _lib.load()
    msg = msg.contiguous()
    d_out = d_out.contiguous()
    dev = msg.device
    E, F = int(msg.shape[0]), int(msg.shape[1])
    N = int(out.shape[0])
    d_msg = torch.empty_like(msg)
    with device_guard(dev):
        check(lib.gcrl_pna_backward(ptr(msg), ptr(out), ptr(var), ptr(d_out), ptr(seg_off), ptr(dst), ptr(amin),
                                    ptr(amax), N, E, F, ptr(d_msg), stream_ptr(dev)), 'gcrl_pna_backward')
    return d_msg


SEGMENT_OPS = {'sum': 0, 'mean': 1, 'min': 2, 'max': 3, 'var': 4, 'std': 5}


def segment_reduce(msg, seg_off, n_vertices, which):
    lib = _lib.load()
    require_cuda(msg, 'msg')
    msg = msg.contiguous()
    dev = msg.device
    E, F = int(msg.shape[0]), int(msg.shape[1])
    out = torch.empty(int(n_vertices), F, dtype=torch.float32, device=dev)
    with device_guard(dev):
        check(lib.gcrl_segment_reduce(ptr(msg), ptr(seg_off), int(n_vertices), E, F, SEGMENT_OPS[which],
                                      ptr(out), stream_ptr(dev)), 'gcrl_segment_reduce')
    return out


# ---- GN network -----------------------------------------------------------------------------
def gn_forward(params, dims, packed, x, eattr, gfeat=None, vgraph=None, with_stash=False,
               math_mode=_lib.GCRL_MATH_FP32, stash=None, ws=None):
    """params: flat fp32 [P]; dims = (vertex_dim, edge_dim, global_dim); x [N,vd]; eattr [E,ed]
    INTERNAL order.  Returns (h [N,64], q [N], stash-or-None)."""
    lib = _lib.load()
    require_cuda(params, 'params')
    dev = params.device
    N, E = packed.n_vertices, packed.n_edges
    x = x.contiguous()
    eattr = eattr.contiguous()
    h = torch.empty(N, HID, dtype=torch.float32, device=dev)
    q = torch.empty(N, dtype=torch.float32, device=dev)
    if with_stash:
        nb = lib.gcrl_gn_stash_bytes(N, E)
        if stash is None or stash.numel() < nb:
            stash = torch.empty(nb, dtype=torch.uint8, device=dev)
    else:
        stash = None
    nb_ws = lib.gcrl_gn_forward_workspace_bytes(N, E, 1 if with_stash else 0)
    if ws is None or ws.numel() < nb_ws:            # ws: a caller-owned workspace (CUDA-graph capture keeps its own)
        ws = workspace(nb_ws, dev, 'gn')
    with device_guard(dev):
        check(lib.gcrl_gn_forward(ptr(params), dims[0], dims[1], dims[2], N, E, ptr(packed.seg_off),
                                  ptr(packed.src), ptr(packed.dst), ptr(x), ptr(eattr), ptr(gfeat), ptr(vgraph),
                                  ptr(stash), ptr(h), ptr(q), ptr(ws), ws.numel(), math_mode, stream_ptr(dev)),
              'gcrl_gn_forward')
    return h, q, stash


def gn_block_forward(params, dims, block, packed, x, eattr, math_mode=_lib.GCRL_MATH_FP32):
    """One GNNBlock: returns (msg_out [E,64] INTERNAL order, upd_out [N,64])."""
    lib = _lib.load()
    dev = params.device
    N, E = packed.n_vertices, packed.n_edges
    x = x.contiguous()
    eattr = eattr.contiguous()
    msg = torch.empty(E, HID, dtype=torch.float32, device=dev)
    upd = torch.empty(N, HID, dtype=torch.float32, device=dev)
    nb = 4 * N * (2 * 64 + 256 + 1) + 2048
    ws = workspace(nb, dev, 'gnblk')
    with device_guard(dev):
        check(lib.gcrl_gn_block_forward(ptr(params), dims[0], dims[1], dims[2], block, N, E, ptr(packed.seg_off),
                                        ptr(packed.src), ptr(packed.dst), ptr(x), ptr(eattr), ptr(msg), ptr(upd),
                                        ptr(ws), ws.numel(), math_mode, stream_ptr(dev)), 'gcrl_gn_block_forward')
    return msg, upd


def gn_backward(params, dims, packed, x, eattr, stash, d_q=None, d_h=None, gfeat=None, vgraph=None,
                grad=None, accumulate=False, math_mode=_lib.GCRL_MATH_FP32, ws=None):
    """Returns the flat gradient [P] (written into `grad` when given)."""
    lib = _lib.load()
    dev = params.device
    N, E = packed.n_vertices, packed.n_edges
    if grad is None:
        grad = torch.empty_like(params)
        accumulate = False
    nb_ws = lib.gcrl_gn_backward_workspace_bytes(N, E)
    if ws is None or ws.numel() < nb_ws:
        ws = workspace(nb_ws, dev, 'gnbwd')
    d_q = d_q.contiguous() if d_q is not None else None
    d_h = d_h.contiguous() if d_h is not None else None
    with device_guard(dev):
        check(lib.gcrl_gn_backward(ptr(params), dims[0], dims[1], dims[2], N, E, ptr(packed.seg_off),
                                   ptr(packed.src), ptr(packed.dst), ptr(x.contiguous()), ptr(eattr.contiguous()),
                                   ptr(gfeat), ptr(vgraph), ptr(stash), ptr(d_q), ptr(d_h), ptr(grad),
                                   1 if accumulate else 0, ptr(ws), ws.numel(), math_mode, stream_ptr(dev)),
              'gcrl_gn_backward')
    return grad


# ---- agent-side scalar kernels ----------------------------------------------------------------
def score_argmax(q, colour, node_off, want_q=False):
    """Per graph: first argmax of q over vertices with colour < 0 (local index; -1 if none)."""
    lib = _lib.load()
    require_cuda(q, 'q')
    dev = q.device
    G = int(node_off.numel()) - 1
    action = _i32(G, dev)
    qb = torch.empty(G, dtype=torch.float32, device=dev) if want_q else None
    with device_guard(dev):
        check(lib.gcrl_score_argmax(ptr(q.contiguous()), ptr(colour), ptr(node_off), G, ptr(action), ptr(qb),
                                    stream_ptr(dev)), 'gcrl_score_argmax')
    return (action, qb) if want_q else action


def td_target_loss(q_cur, q_next, next_colour, node_off, action, reward, done, gamma, batch_total=None,
                   loss_out=None, err=None):
    """Returns (loss [1], q_expected [G], q_target [G], d_q [N]).  loss_out: a [1] tensor the loss is ADDED to
    (micro-batches accumulate into one scalar); err: int32 [1] set to 1 + g when graph g's action is outside
    [0, n_g) (that graph then contributes nothing instead of touching memory out of bounds)."""
    lib = _lib.load()
    dev = q_cur.device
    G = int(node_off.numel()) - 1
    if batch_total is None:
        batch_total = G
    q_exp = torch.empty(G, dtype=torch.float32, device=dev)
    q_tgt = torch.empty(G, dtype=torch.float32, device=dev)
    d_q = torch.empty_like(q_cur)
    loss = loss_out if loss_out is not None else torch.zeros(1, dtype=torch.float32, device=dev)
    with device_guard(dev):
        check(lib.gcrl_td_target_loss(ptr(q_cur.contiguous()), ptr(q_next.contiguous()), ptr(next_colour),
                                      ptr(node_off), G, ptr(action), ptr(reward), ptr(done), float(gamma),
                                      int(batch_total), ptr(q_exp), ptr(q_tgt), ptr(d_q), ptr(loss), ptr(err),
                                      stream_ptr(dev)), 'gcrl_td_target_loss')
    return loss, q_exp, q_tgt, d_q


def adam_soft_update(params, grad, exp_avg, exp_avg_sq, target, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                     tau=1e-3):
    lib = _lib.load()
    dev = params.device
    with device_guard(dev):
        check(lib.gcrl_adam_soft_update(ptr(params), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), ptr(target),
                                        params.numel(), float(lr), float(betas[0]), float(betas[1]), float(eps),
                                        int(step), float(tau), stream_ptr(dev)), 'gcrl_adam_soft_update')


def soft_update(behaviour, target, tau):
    lib = _lib.load()
    dev = behaviour.device
    with device_guard(dev):
        check(lib.gcrl_soft_update(ptr(behaviour), ptr(target), behaviour.numel(), float(tau), stream_ptr(dev)),
              'gcrl_soft_update')


# ---- environment ----------------------------------------------------------------------------------
def env_reset(colour, max_colour, x, node_off):
    lib = _lib.load()
    dev = colour.device
    G = int(node_off.numel()) - 1
    with device_guard(dev):
        check(lib.gcrl_env_reset(ptr(colour), ptr(max_colour), ptr(x), ptr(node_off), G, colour.numel(),
                                 stream_ptr(dev)), 'gcrl_env_reset')


def env_step(adj, adj_off, node_off, action, colour, max_colour, reward, done, x):
    lib = _lib.load()
    dev = colour.device
    G = int(node_off.numel()) - 1
    with device_guard(dev):
        check(lib.gcrl_env_step(ptr(adj), ptr(adj_off), ptr(node_off), G, ptr(action), ptr(colour),
                                ptr(max_colour), ptr(reward), ptr(done), ptr(x), stream_ptr(dev)), 'gcrl_env_step')


def env_step_greedy(adj, adj_off, node_off, q, colour, max_colour, reward, done, x, action_out=None):
    """Masked per-graph first-argmax of q fused with the environment step; returns the actions taken (int32 [G])."""
    lib = _lib.load()
    dev = colour.device
    G = int(node_off.numel()) - 1
    if action_out is None:
        action_out = _i32(G, dev)
    with device_guard(dev):
        check(lib.gcrl_env_step_greedy(ptr(adj), ptr(adj_off), ptr(node_off), G, ptr(q.contiguous()), ptr(action_out), ptr(colour),
                                       ptr(max_colour), ptr(reward), ptr(done), ptr(x), stream_ptr(dev)), 'gcrl_env_step_greedy')
    return action_out


def env_first_vertex(adj, adj_off, node_off, n_vertices, want_counts=False):
    lib = _lib.load()
    dev = adj.device
    G = int(node_off.numel()) - 1
    action = _i32(G, dev)
    counts = torch.empty(int(n_vertices), 3, dtype=torch.int32, device=dev) if want_counts else None
    with device_guard(dev):
        check(lib.gcrl_env_first_vertex(ptr(adj), ptr(adj_off), ptr(node_off), G, ptr(action), ptr(counts),
                                        stream_ptr(dev)), 'gcrl_env_first_vertex')
    return (action, counts) if want_counts else action


# ---- device replay ring ------------------------------------------------------------------------------
def replay_store(slot, graph_id, node_off, cur_colour, next_colour, action, reward, done, ring):
    """ring = (r_graph, r_cur [C,stride] int16, r_next, r_action, r_reward, r_done); slot int32 [n_envs], < 0 skips."""
    lib = _lib.load()
    require_cuda(cur_colour, 'cur_colour')
    dev = cur_colour.device
    r_graph, r_cur, r_next, r_action, r_reward, r_done = ring
    with device_guard(dev):
        check(lib.gcrl_replay_store(int(slot.numel()), ptr(slot), ptr(graph_id), ptr(node_off), ptr(cur_colour),
                                    ptr(next_colour), ptr(action), ptr(reward), ptr(done), int(r_cur.shape[1]),
                                    ptr(r_graph), ptr(r_cur), ptr(r_next), ptr(r_action), ptr(r_reward), ptr(r_done),
                                    stream_ptr(dev)), 'gcrl_replay_store')


def replay_gather(slots, ring, pool_adj, pool_adj_off, pool_n, n_max, out_node_off, out_adj_off, n_vertices,
                  n_adj_bytes, err=None):
    """-> (adj uint8 [sum n^2], cur int32 [N], next int32 [N], action int32 [B], reward f32 [B], done int32 [B])."""
    lib = _lib.load()
    require_cuda(slots, 'slots')
    dev = slots.device
    r_graph, r_cur, r_next, r_action, r_reward, r_done = ring
    B = int(slots.numel())
    adj = torch.empty(int(n_adj_bytes), dtype=torch.uint8, device=dev)
    cur, nxt = _i32(n_vertices, dev), _i32(n_vertices, dev)
    action, done = _i32(B, dev), _i32(B, dev)
    reward = torch.empty(B, dtype=torch.float32, device=dev)
    with device_guard(dev):
        check(lib.gcrl_replay_gather(B, ptr(slots), ptr(r_graph), ptr(r_cur), ptr(r_next), ptr(r_action),
                                     ptr(r_reward), ptr(r_done), int(r_cur.shape[1]), ptr(pool_adj),
                                     ptr(pool_adj_off), ptr(pool_n), int(n_max), ptr(out_node_off),
                                     ptr(out_adj_off), ptr(adj), ptr(cur), ptr(nxt), ptr(action), ptr(reward),
                                     ptr(done), ptr(err), stream_ptr(dev)), 'gcrl_replay_gather')
    return adj, cur, nxt, action, reward, done
