// segreduce.cuh -- CTA-level segmented PNA reduction of a 128-row message tile held in
// shared memory (rows = consecutive destination-major edge slots).  Shared by the fused
// edge kernels; deterministic (fixed combination order, no atomics).
//
// 256 threads: col = tid & 63, quarter q = tid >> 6 scans rows [32q, 32q+32) of its column.
// A quarter emits a "head" partial (rows before its first segment boundary, possibly the
// continuation of an earlier quarter / tile), finalises segments that start and end inside
// it, and emits a "tail" partial (rows after its last boundary).  Threads 0..63 then chain
// carry (+) head / tail across the four quarters; the carry lives in their registers across
// the tiles of a work item.
#pragma once
#include "common.cuh"

namespace gcrl {

struct Stat {
    float sum, sq, mn, mx;
    int32_t amn, amx;
};
__device__ __forceinline__ void stat_init(Stat& s) {
    s.sum = 0.f; s.sq = 0.f; s.mn = INFINITY; s.mx = -INFINITY;
    s.amn = 0x7fffffff; s.amx = 0x7fffffff;
}
// rows fed in ascending slot order; strict comparisons keep the FIRST extremum
// (torch_scatter CPU kernel: `if (src < out) {out = src; arg = e}` in edge order).
__device__ __forceinline__ void stat_add(Stat& s, float v, int32_t slot) {
    s.sum += v;
    s.sq = fmaf(v, v, s.sq);
    if (v < s.mn) { s.mn = v; s.amn = slot; }
    if (v > s.mx) { s.mx = v; s.amx = slot; }
}
__device__ __forceinline__ void stat_merge(Stat& a, const Stat& b) {
    a.sum += b.sum;
    a.sq += b.sq;
    if (b.mn < a.mn || (b.mn == a.mn && b.amn < a.amn)) { a.mn = b.mn; a.amn = b.amn; }
    if (b.mx > a.mx || (b.mx == a.mx && b.amx < a.amx)) { a.mx = b.mx; a.amx = b.amx; }
}
// principal_neighbourhood_aggregation.py:45-46,57-64.  Non-contracted mul/sub: the variance
// is cancellation-prone and the reference rounds mean*mean and the subtraction separately.
__device__ __forceinline__ void stat_final(const Stat& s, int32_t deg, float& mean, float& mn,
                                           float& mx, float& sd, float& var) {
    if (deg <= 0) {  // empty segment: scatter leaves zeros
        mean = 0.f; mn = 0.f; mx = 0.f; var = 0.f; sd = sqrtf(PNA_EPS);
        return;
    }
    float cnt = (float)deg;
    mean = __fdiv_rn(s.sum, cnt);
    float msq = __fdiv_rn(s.sq, cnt);
    var = __fsub_rn(msq, __fmul_rn(mean, mean));
    sd = sqrtf(__fadd_rn(fmaxf(var, 0.f), PNA_EPS));
    mn = s.mn; mx = s.mx;
}

// value-only variant (inference: nobody reads the arg slots)
__device__ __forceinline__ void stat_add_noarg(Stat& s, float v) {
    s.sum += v;
    s.sq = fmaf(v, v, s.sq);
    s.mn = fminf(s.mn, v);
    s.mx = fmaxf(s.mx, v);
}

struct AggOut {
    float* agg;       // [N,256] mean|min|max|std
    float* var;       // [N,64] raw variance (relu' mask for backward) or null
    int32_t* amin;    // [N,64] or null
    int32_t* amax;    // [N,64] or null
    const int32_t* seg_off;
};

__device__ __forceinline__ void agg_write(const AggOut& o, int32_t v, int col, const Stat& s) {
    int32_t deg = o.seg_off[v + 1] - o.seg_off[v];
    float mean, mn, mx, sd, var;
    stat_final(s, deg, mean, mn, mx, sd, var);
    float* a = o.agg + (int64_t)v * 256;
    a[col] = mean; a[64 + col] = mn; a[128 + col] = mx; a[192 + col] = sd;
    if (o.var) o.var[(int64_t)v * 64 + col] = var;
    if (o.amin) o.amin[(int64_t)v * 64 + col] = s.amn;
    if (o.amax) o.amax[(int64_t)v * 64 + col] = s.amx;
}
// out-of-line copy for the warp-specialised kernels, whose hot loops have to fit the instruction cache (it runs
// once per (segment, column) but is referenced from several places)
static __device__ __noinline__ void agg_write_ool(const AggOut& o, int32_t v, int col, const Stat& s) { agg_write(o, v, col, s); }

// scratch carved from shared memory: 2 (head/tail) x 6 fields x 4 quarters x 64 cols + ids
struct RedScratch {
    float f[2][4][4][64];     // [head/tail][sum,sq,mn,mx][quarter][col]
    int32_t a[2][2][4][64];   // [head/tail][amn,amx][quarter][col]
    int32_t head_dst[4];
    int32_t tail_dst[4];      // -1: quarter saw no boundary (everything is in head)
};

struct Carry {
    Stat s;
    int32_t dst;  // -1 = empty
};

// Phase A: all 256 threads.  load(r, col) reads element (r, col) of the 128-row message tile in
// shared memory, dst_s[128] destination ids of rows.
template <class LoadFn>
__device__ __forceinline__ void tile_reduce_scan_fn(LoadFn load, const int32_t* dst_s, int nrows,
                                                    int32_t slot0, RedScratch* sc, const AggOut& out) {
    const int col = threadIdx.x & 63, q = threadIdx.x >> 6;
    const int r0 = q * 32;
    const int r1 = min(r0 + 32, nrows);
    if (r0 >= r1) return;
    Stat cur;
    stat_init(cur);
    int32_t cd = dst_s[r0];
    const int32_t hd = cd;
    bool seen = false;
    for (int r = r0; r < r1; ++r) {
        int32_t d = dst_s[r];
        if (d != cd) {
            if (!seen) {
                sc->f[0][0][q][col] = cur.sum; sc->f[0][1][q][col] = cur.sq;
                sc->f[0][2][q][col] = cur.mn;  sc->f[0][3][q][col] = cur.mx;
                sc->a[0][0][q][col] = cur.amn; sc->a[0][1][q][col] = cur.amx;
                seen = true;
            } else {
                agg_write(out, cd, col, cur);   // segment lies wholly inside this quarter
            }
            stat_init(cur);
            cd = d;
        }
        stat_add(cur, load(r, col), slot0 + r);
    }
    const int w = seen ? 1 : 0;
    sc->f[w][0][q][col] = cur.sum; sc->f[w][1][q][col] = cur.sq;
    sc->f[w][2][q][col] = cur.mn;  sc->f[w][3][q][col] = cur.mx;
    sc->a[w][0][q][col] = cur.amn; sc->a[w][1][q][col] = cur.amx;
    if (col == 0) {
        sc->head_dst[q] = hd;
        sc->tail_dst[q] = seen ? cd : -1;
    }
}

// Phase A for the SW128-slab fp32 tile of the tensor-core kernels (tc.cuh): rows are walked in
// groups of 8 so the XOR swizzle (chunk ^ (r & 7)) folds into eight per-thread constant offsets;
// the 8 values of a group are loaded up front and segment boundaries come from a per-tile bit
// mask (bit r set = row r starts a new destination), so no load sits on the branch chain.
template <bool WANT_ARG>
__device__ __forceinline__ void tile_reduce_scan_sw_t(const float* L, const int32_t* dst_s, const uint32_t* start_bits,
                                                      int nrows, int32_t slot0, RedScratch* sc, const AggOut& out, int ltid);
template <bool WANT_ARG>
__device__ __forceinline__ void tile_reduce_scan_sw(const float* L, const int32_t* dst_s, const uint32_t* start_bits,
                                                    int nrows, int32_t slot0, RedScratch* sc, const AggOut& out) {
    tile_reduce_scan_sw_t<WANT_ARG>(L, dst_s, start_bits, nrows, slot0, sc, out, (int)threadIdx.x);
}
// ltid: index of the thread inside the 256-thread reducer group
template <bool WANT_ARG>
__device__ __forceinline__ void tile_reduce_scan_sw_t(const float* L, const int32_t* dst_s, const uint32_t* start_bits,
                                                      int nrows, int32_t slot0, RedScratch* sc, const AggOut& out, int ltid) {
    const int col = ltid & 63, q = ltid >> 6;
    const int r0 = q * 32;
    const int r1 = min(r0 + 32, nrows);
    if (r0 >= r1) return;
    const int c4 = col >> 2, cc = c4 & 7;
    const unsigned char* base = reinterpret_cast<const unsigned char*>(L) + (c4 >> 3) * (128 * 128) + (col & 3) * 4;
    const uint32_t starts = start_bits[q] & ~1u;    // a boundary at the quarter's first row is the head itself
    const int32_t hd = dst_s[r0];
    const int nstart = __popc(starts);
    if (nstart <= 1) {
        // at most one segment boundary inside the quarter (always, for graphs of >= 34 vertices): rows [r0, ra) feed
        // the head partial A, rows [ra, r1) the tail partial B; groups of 8 rows wholly on one side run straight-line
        const int bl = nstart ? (__ffs(starts) - 1) : 32;
        const int ra = min(r0 + bl, r1);
        Stat A, B;
        stat_init(A);
        stat_init(B);
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
            const int rb = r0 + 8 * g;
            if (rb >= r1) break;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = *reinterpret_cast<const float*>(base + (rb + k) * 128 + ((cc ^ k) << 4));
            if (rb + 8 <= ra) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (WANT_ARG) stat_add(A, v[k], slot0 + rb + k);
                    else stat_add_noarg(A, v[k]);
                }
            } else if (rb >= ra && rb + 8 <= r1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (WANT_ARG) stat_add(B, v[k], slot0 + rb + k);
                    else stat_add_noarg(B, v[k]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int r = rb + k;
                    if (r < ra) {
                        if (WANT_ARG) stat_add(A, v[k], slot0 + r);
                        else stat_add_noarg(A, v[k]);
                    } else if (r < r1) {
                        if (WANT_ARG) stat_add(B, v[k], slot0 + r);
                        else stat_add_noarg(B, v[k]);
                    }
                }
            }
        }
        sc->f[0][0][q][col] = A.sum; sc->f[0][1][q][col] = A.sq;
        sc->f[0][2][q][col] = A.mn;  sc->f[0][3][q][col] = A.mx;
        sc->a[0][0][q][col] = A.amn; sc->a[0][1][q][col] = A.amx;
        const bool has_tail = nstart == 1 && r0 + bl < r1;
        if (has_tail) {
            sc->f[1][0][q][col] = B.sum; sc->f[1][1][q][col] = B.sq;
            sc->f[1][2][q][col] = B.mn;  sc->f[1][3][q][col] = B.mx;
            sc->a[1][0][q][col] = B.amn; sc->a[1][1][q][col] = B.amx;
        }
        if (col == 0) {
            sc->head_dst[q] = hd;
            sc->tail_dst[q] = has_tail ? dst_s[r0 + bl] : -1;
        }
        return;
    }
    Stat cur;
    stat_init(cur);
    int32_t cd = hd;
    bool seen = false;
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        const int rb = r0 + 8 * g;
        if (rb >= r1) break;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = *reinterpret_cast<const float*>(base + (rb + k) * 128 + ((cc ^ k) << 4));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = rb + k;
            if (r < r1) {
                if ((starts >> (8 * g + k)) & 1u) {
                    if (!seen) {
                        sc->f[0][0][q][col] = cur.sum; sc->f[0][1][q][col] = cur.sq;
                        sc->f[0][2][q][col] = cur.mn;  sc->f[0][3][q][col] = cur.mx;
                        sc->a[0][0][q][col] = cur.amn; sc->a[0][1][q][col] = cur.amx;
                        seen = true;
                    } else {
                        agg_write(out, cd, col, cur);
                    }
                    stat_init(cur);
                    cd = dst_s[r];
                }
                if (WANT_ARG) stat_add(cur, v[k], slot0 + r);
                else stat_add_noarg(cur, v[k]);
            }
        }
    }
    const int w = seen ? 1 : 0;
    sc->f[w][0][q][col] = cur.sum; sc->f[w][1][q][col] = cur.sq;
    sc->f[w][2][q][col] = cur.mn;  sc->f[w][3][q][col] = cur.mx;
    sc->a[w][0][q][col] = cur.amn; sc->a[w][1][q][col] = cur.amx;
    if (col == 0) {
        sc->head_dst[q] = hd;
        sc->tail_dst[q] = seen ? cd : -1;
    }
}

// M_s: tile [128][ld] row-major in smem
__device__ __forceinline__ void tile_reduce_scan(const float* M_s, int ld, const int32_t* dst_s,
                                                 int nrows, int32_t slot0, RedScratch* sc,
                                                 const AggOut& out) {
    tile_reduce_scan_fn([=](int r, int col) { return M_s[r * ld + col]; }, dst_s, nrows, slot0, sc, out);
}

// Phase B: threads 0..63 only (after a __syncthreads following phase A).
__device__ __forceinline__ void tile_reduce_chain_t(int nrows, const RedScratch* sc, Carry& c, const AggOut& out, int col);
__device__ __forceinline__ void tile_reduce_chain(int nrows, const RedScratch* sc, Carry& c,
                                                  const AggOut& out) {
    tile_reduce_chain_t(nrows, sc, c, out, (int)threadIdx.x);
}
__device__ __forceinline__ void tile_reduce_chain_t(int nrows, const RedScratch* sc, Carry& c, const AggOut& out, int col) {
    for (int q = 0; q < 4; ++q) {
        if (q * 32 >= nrows) break;
        Stat h;
        h.sum = sc->f[0][0][q][col]; h.sq = sc->f[0][1][q][col];
        h.mn = sc->f[0][2][q][col];  h.mx = sc->f[0][3][q][col];
        h.amn = sc->a[0][0][q][col]; h.amx = sc->a[0][1][q][col];
        const int32_t hd = sc->head_dst[q];
        if (c.dst == hd) {
            stat_merge(c.s, h);
        } else {
            if (c.dst >= 0) agg_write(out, c.dst, col, c.s);
            c.s = h; c.dst = hd;
        }
        const int32_t td = sc->tail_dst[q];
        if (td >= 0) {
            agg_write(out, c.dst, col, c.s);
            c.s.sum = sc->f[1][0][q][col]; c.s.sq = sc->f[1][1][q][col];
            c.s.mn = sc->f[1][2][q][col];  c.s.mx = sc->f[1][3][q][col];
            c.s.amn = sc->a[1][0][q][col]; c.s.amx = sc->a[1][1][q][col];
            c.dst = td;
        }
    }
}

__device__ __forceinline__ void carry_flush_t(Carry& c, const AggOut& out, int col) {
    if (c.dst >= 0) agg_write(out, c.dst, col, c.s);
    c.dst = -1;
}
__device__ __forceinline__ void carry_flush(Carry& c, const AggOut& out) { carry_flush_t(c, out, (int)threadIdx.x); }

// lower_bound over seg_off[0..N]: first v in [0,N] with seg_off[v] >= key, clamped to N
__device__ __forceinline__ int32_t seg_lower_bound(const int32_t* __restrict__ seg_off, int32_t N,
                                                   int64_t key) {
    int32_t lo = 0, hi = N;  // answer in [lo, hi]
    while (lo < hi) {
        int32_t mid = (lo + hi) >> 1;
        if ((int64_t)seg_off[mid] >= key) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Warp-cooperative flavour for two keys at once: 32 probes per key and round, so a 2 000-vertex batch takes 3 rounds
// of (two overlapped) L2 round trips instead of 2 x 11 dependent ones -- with work items of a single tile (small
// batches) the serial search was a third of the item's time.  All lanes return the same values.
__device__ __forceinline__ void seg_lower_bound2_warp(const int32_t* __restrict__ seg_off, int32_t N, int64_t key0,
                                                      int64_t key1, int lane, int32_t& out0, int32_t& out1) {
    int32_t lo0 = 0, hi0 = N, lo1 = 0, hi1 = N;      // answers in [lo, hi]
    while (lo0 < hi0 || lo1 < hi1) {
        const int32_t st0 = (hi0 - lo0 + 31) >> 5, st1 = (hi1 - lo1 + 31) >> 5;
        const int32_t p0 = lo0 + lane * st0, p1 = lo1 + lane * st1;
        const bool a0 = lo0 < hi0, a1 = lo1 < hi1;
        int32_t s0 = 0, s1 = 0;
        if (a0 && p0 < hi0) s0 = __ldg(seg_off + p0);
        if (a1 && p1 < hi1) s1 = __ldg(seg_off + p1);
        const uint32_t m0 = __ballot_sync(0xffffffffu, p0 >= hi0 || (int64_t)s0 >= key0);
        const uint32_t m1 = __ballot_sync(0xffffffffu, p1 >= hi1 || (int64_t)s1 >= key1);
        if (a0) {
            if (m0 == 0u) lo0 = lo0 + 31 * st0 + 1;
            else {
                const int f = __ffs(m0) - 1;
                const int32_t nh = lo0 + f * st0;
                if (f > 0) lo0 = lo0 + (f - 1) * st0 + 1;
                hi0 = nh < hi0 ? nh : hi0;
            }
        }
        if (a1) {
            if (m1 == 0u) lo1 = lo1 + 31 * st1 + 1;
            else {
                const int f = __ffs(m1) - 1;
                const int32_t nh = lo1 + f * st1;
                if (f > 0) lo1 = lo1 + (f - 1) * st1 + 1;
                hi1 = nh < hi1 ? nh : hi1;
            }
        }
    }
    out0 = lo0;
    out1 = lo1;
}

}  // namespace gcrl
