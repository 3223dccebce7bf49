"""Drop-in for the reference's principal_neighbourhood_aggregation.py: the AGGREGATORS and
SCALERS tables with the same names and call signatures
(`AGGREGATORS[name](src[E,F], index[E] int64, dim_size) -> [dim_size,F]`,
`SCALERS[name](src, deg, avg_deg)`), computed by the segment kernels of libgcrl_b200
(gcrl_pack + gcrl_segment_reduce / gcrl_pna_forward) instead of torch_scatter.

Reference: principal_neighbourhood_aggregation.py:6-38 (scalers), :41-64 (aggregators),
:67-74 (tables).  CUDA tensors only -- there is no CPU path.
"""
import torch

from . import ops


def _segments(index, dim_size):
    """Destination-major layout of an index vector (torch_scatter semantics: dim_size=None
    means max(index)+1)."""
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    ei = torch.stack([index, index], 0)
    return ops.pack(ei, dim_size), int(dim_size)


class _SegmentReduce(torch.autograd.Function):
    """One reduction with torch_scatter's gradient routing (min/max: first arg edge only)."""

    @staticmethod
    def forward(ctx, src, index, dim_size, which):
        P, N = _segments(index, dim_size)
        src_i = ops.permute_rows(src, P.perm, True)
        F = src.shape[1]
        out4, amin, amax, var = ops.pna_forward(src_i, P.seg_off, N, want_backward_state=True)
        if which == 'sum':
            out = ops.segment_reduce(src_i, P.seg_off, N, 'sum')
        elif which == 'var':
            out = var
        else:
            col = {'mean': 0, 'min': 1, 'max': 2, 'std': 3}[which]
            out = out4[:, col * F:(col + 1) * F].contiguous()
        ctx.which, ctx.P, ctx.N = which, P, N
        ctx.save_for_backward(src_i, out4, amin, amax, var)
        return out

    @staticmethod
    def backward(ctx, g):
        src_i, out4, amin, amax, var = ctx.saved_tensors
        P, N, which = ctx.P, ctx.N, ctx.which
        F = src_i.shape[1]
        g = g.contiguous()
        d4 = torch.zeros(N, 4 * F, dtype=torch.float32, device=g.device)
        dst = P.dst.long()
        if which == 'sum':
            d_i = g[dst]
        elif which == 'var':
            # var = mean(x^2) - mean(x)^2  ->  d/dx = 2 (x - mean) / deg
            deg = (P.seg_off[1:] - P.seg_off[:-1]).clamp(min=1).float().unsqueeze(1)
            d_i = (g / deg)[dst] * 2.0 * (src_i - out4[:, :F][dst])
        else:
            col = {'mean': 0, 'min': 1, 'max': 2, 'std': 3}[which]
            d4[:, col * F:(col + 1) * F] = g
            d_i = ops.pna_backward(src_i, out4, var, d4, P.seg_off, P.dst, amin, amax)
        return ops.permute_rows(d_i, P.perm, False), None, None, None


def aggregate_sum(src, index, dim_size):
    return _SegmentReduce.apply(src, index, dim_size, 'sum')


def aggregate_mean(src, index, dim_size):
    return _SegmentReduce.apply(src, index, dim_size, 'mean')


def aggregate_min(src, index, dim_size):
    return _SegmentReduce.apply(src, index, dim_size, 'min')


def aggregate_max(src, index, dim_size):
    return _SegmentReduce.apply(src, index, dim_size, 'max')


def aggregate_var(src, index, dim_size):
    return _SegmentReduce.apply(src, index, dim_size, 'var')


def aggregate_std(src, index, dim_size):
    return _SegmentReduce.apply(src, index, dim_size, 'std')


AGGREGATORS = {
    'sum': aggregate_sum,
    'mean': aggregate_mean,
    'min': aggregate_min,
    'max': aggregate_max,
    'var': aggregate_var,
    'std': aggregate_std,
}


# ---- degree scalers (elementwise; selected by name, the GN uses 'identity' only) ----------
def scale_identity(src, deg, avg_deg):
    return src


def scale_amplification(src, deg, avg_deg):
    return src * (torch.log(deg + 1) / avg_deg['log'])


def scale_attenuation(src, deg, avg_deg):
    factor = torch.where(deg == 0, torch.ones_like(deg), avg_deg['log'] / torch.log(deg + 1))
    return src * factor


def scale_linear(src, deg, avg_deg):
    return src * (deg / avg_deg['lin'])


def scale_inverse_linear(src, deg, avg_deg):
    factor = torch.where(deg == 0, torch.ones_like(deg), avg_deg['lin'] / deg)
    return src * factor


SCALERS = {
    'identity': scale_identity,
    'amplification': scale_amplification,
    'attenuation': scale_attenuation,
    'linear': scale_linear,
    'inverse_linear': scale_inverse_linear,
}
