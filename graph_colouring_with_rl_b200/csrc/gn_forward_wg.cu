// gn_forward_wg.cu -- fused edge kernel of the GN forward pass, four tiles in flight per SM.
//
// Reference: architecture.py:51-66.  Message MLP on tcgen05 (bf16 / bf16x3 split operands, fp32 accumulate in TMEM),
// e_l stored by TMA, PNA aggregation fused in.  Round 1's first design (removed; git history) ran the
// phases one after the other in a 256-thread CTA and relied on a second CTA per SM to fill the gaps; its counters
// said latency: 40 % issue utilisation, a quarter of the stall samples on barriers and
// another quarter on loads, 16 warps per SM.  Shared memory (110 KB per CTA, 32 KB of it a private copy of the
// weights) is what stops a third CTA.  This kernel keeps the phase structure but packs FOUR independent 128-thread
// warp groups into one CTA per SM:
//   * the weight tiles (32 KB) are shared by the four groups;
//   * a group owns ONE 32 KB tile buffer T that is, in turn, the TMA landing tile of e_{l-1} (fp32), the bf16 hi/lo
//     operand tile (converted in place through registers), the hid operand tile and the fp32 staging tile of m_l
//     (TMA store source + reducer input);
//   * Pi[dst] + Pj[src] never touch shared memory: they are gathered in the 16x256b TMEM fragment layout and
//     written with tcgen05.st, G1 accumulates on top (as in the backward, gn_backward_wg.cu);
//   * groups synchronise with named barriers and their own mbarriers only, so one group's MMA / TMA / gather
//     latency is covered by the other three.
// TMEM: 128 columns per group (acc1, acc2) = all 512.  Shared memory: 32 + 4 x (32 + 12) + ids = 214 KB.
// Late round 2 (the default, template parameter ATM): the A operands of both products live in TENSOR memory -- e_{l-1}
// is split into the columns of acc2 (idle until G2), relu(z1) is written back in place over acc1 -- so T is only the
// TMA landing tile and the fp32 staging tile of m_l; block 1 with the reference's input widths (DIRECT) evaluates z1
// from five scalars per edge instead of gathering projections.  DESIGN.md 4.2a / 4.2b.
#include "gn_common.cuh"
#include "ws.cuh"
#include <stdlib.h>

namespace gcrl {

using namespace tc;

// phase timeline of group 0 of CTA 0 (diagnostic build: GCRL_NVCC_FLAGS=-DGCRL_FWD_TRACE_BUILD)
#ifdef GCRL_FWD_TRACE_BUILD
#define FW_TR(k) do { if (tid == 0 && blockIdx.x == 0) { const long long c_ = clock64(); tr_acc[k] += c_ - tr_last; tr_last = c_; } } while (0)
#else
#define FW_TR(k) do { } while (0)
#endif

constexpr int WG_N = 4;
constexpr int WG_THREADS = 128 * WG_N;

struct FwdWgSmem {
    __nv_bfloat16 Wc[2][64 * 64];        // C hi/lo (FIRST: C^T in fp32 [d][o])
    __nv_bfloat16 W2[2][64 * 64];
    unsigned char T[WG_N][32768];        // per group: landing fp32 -> operand hi|lo -> hid hi|lo -> m fp32
    RedScratch red[WG_N];
    float b2[64];
    int32_t dst[WG_N][TILE];
    int32_t src[WG_N][TILE];
    uint32_t start_bits[WG_N][4];
    uint64_t bar_tma[WG_N], bar_mma[WG_N];
    uint32_t tmem_base;
};

struct FwdWgArgs {
    const float* Pi; const float* Pj;
    const float* e_prev; int De;
    const float* W1; int ldW1, coff;
    const float* W2; const float* b2;
    const float* x; const float* b1;          // DIRECT (block 1 with 2 vertex features + 1 edge feature): raw inputs
    const int32_t* seg_off; const int32_t* src; const int32_t* dst;
    int32_t N, E;
    float* m_out;
    AggOut out;
    int item_edges;                           // target edges per work item (an item is a run of whole segments)
    const unsigned char* img_w2; const unsigned char* img_wc;   // pre-converted weight tiles (weight_image.cu) or null
};

// DIRECT (FIRST only; Dv = 2, De = 1 -- the reference's block 1): z1 depends on five scalars per edge, so the row's thread
// evaluates it straight from x[dst], x[src] and the edge feature (the operation order of k_node_proj_first + the
// gather path, bit for bit) and writes relu(z1) into the operand tile: no Pi / Pj rows gathered through L2, no TMEM
// round trip, no projection launch.  The gather phase was 5.4 k of the 18.7 k cycles of a block-1 tile.
// ATM: the A operands of both products live in TENSOR memory instead of shared memory (tcgen05.mma with A read from
// TMEM): e_{l-1} goes landing tile -> registers -> bf16 hi | lo columns over the (still unused) accumulator of G2, and
// relu(z1) is written back IN PLACE over the accumulator columns it was read from, 32 at a time.  Per tile that removes
// 64 KB of operand writes and ~96 KB of MMA operand reads from shared memory (of ~380 KB), the named barrier of the
// in-place conversion and the generic->async proxy fences in front of both products.
template <bool FIRST, bool X3, bool ARG, bool DIRECT = false, bool ATM = false>
__global__ void __launch_bounds__(WG_THREADS, 1) k_edge_fwd_wg(const __grid_constant__ CUtensorMap tm_in,
                                                              const __grid_constant__ CUtensorMap tm_out, const FwdWgArgs a) {
    pdl_prologue();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    FwdWgSmem& S = *reinterpret_cast<FwdWgSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, wg = tid >> 7, lt = tid & 127, w = lt >> 5, lane = tid & 31;
    const int bar_id = 1 + wg;                     // named barrier of this group (0 is __syncthreads)

    if (tid == 0) {
        for (int g = 0; g < WG_N; ++g) { mbar_init(&S.bar_tma[g], 1); mbar_init(&S.bar_mma[g], 1); }
        fence_barrier_init();
        if (!FIRST) tma_prefetch_desc(&tm_in);
        if (a.m_out) tma_prefetch_desc(&tm_out);
    }
    if (tid < 32) tmem_alloc<512>(&S.tmem_base);
    if (a.img_w2) copy_image(S.W2, a.img_w2, 16384, tid, WG_THREADS);
    else load_weight_tile(a.W2, 64, 0, S.W2[0], S.W2[1], tid, WG_THREADS);
    if (!FIRST) {
        if (a.img_wc) copy_image(S.Wc, a.img_wc, 16384, tid, WG_THREADS);
        else load_weight_tile(a.W1, a.ldW1, a.coff, S.Wc[0], S.Wc[1], tid, WG_THREADS);
    }
    float (*Cf)[64] = reinterpret_cast<float (*)[64]>(&S.Wc[0][0]);   // FIRST: C^T in fp32 [d][o]
    // DIRECT: per output column o, (A[o][0], A[o][1], B[o][0], B[o][1]) and (C[o][0], b1[o])
    float4* const wA = reinterpret_cast<float4*>(&S.Wc[0][0]);
    float2* const wC = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(&S.Wc[0][0]) + 1024);
    if (FIRST && !DIRECT)
        for (int t = tid; t < a.De * 64; t += WG_THREADS) Cf[t >> 6][t & 63] = a.W1[(t & 63) * a.ldW1 + a.coff + (t >> 6)];
    if (DIRECT && tid < 64) {
        const float* w = a.W1 + tid * a.ldW1;
        wA[tid] = make_float4(w[0], w[1], w[2], w[3]);
        wC[tid] = make_float2(w[4], a.b1[tid]);
    }
    if (tid < 64) S.b2[tid] = a.b2[tid];
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t t_acc1 = S.tmem_base + 128 * wg, t_acc2 = t_acc1 + 64;
    const uint32_t lane_addr = (uint32_t)(32 * w) << 16;
    const uint32_t idesc = make_idesc_bf16(128, 64, false, false);
    unsigned char* const T = S.T[wg];
    float* const M = reinterpret_cast<float*>(T);
    __nv_bfloat16* const Bh = reinterpret_cast<__nv_bfloat16*>(T);
    __nv_bfloat16* const Bl = reinterpret_cast<__nv_bfloat16*>(T + 16384);
    int32_t* const dst_s = S.dst[wg];
    int32_t* const src_s = S.src[wg];
    RedScratch* const red = &S.red[wg];
    uint64_t* const bar_tma = &S.bar_tma[wg];
    uint64_t* const bar_mma = &S.bar_mma[wg];
    uint32_t ph_tma = 0, ph_mma = 0;
    const int cp = 2 * (lane & 3);
    int32_t cur_d = -1;                            // destination whose Pi row is cached in piv
    float2 piv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) piv[c] = make_float2(0.f, 0.f);

#ifdef GCRL_FWD_TRACE_BUILD
    long long tr_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tr_last = clock64();
    int tr_tiles = 0;
#endif
    const int64_t IE = a.item_edges;
    const int64_t n_items = ((int64_t)a.E + IE - 1) / IE;
    for (int64_t it = (int64_t)blockIdx.x * WG_N + wg; it < n_items; it += (int64_t)gridDim.x * WG_N) {
        int32_t v0, v1;
        seg_lower_bound2_warp(a.seg_off, a.N, it * IE, (it + 1) * IE, lane, v0, v1);
        if (it + 1 == n_items) v1 = a.N;
        const int32_t e0 = a.seg_off[v0], e1 = a.seg_off[v1];
        Carry carry;
        carry.dst = -1;
        stat_init(carry.s);
        bool chain_pending = false;                // the previous tile's partials still wait in the reducer scratch
        int prev_nrows = 0;
        if (!FIRST && lt == 0 && e0 < e1) {
            mbar_arrive_expect_tx(bar_tma, 128 * 64 * 4);
            tma_load_2d(T, &tm_in, 0, e0, bar_tma);
            tma_load_2d(T + 16384, &tm_in, 32, e0, bar_tma);
        }
        int32_t nd = -1, ns = 0;
        if (e0 + lt < e1) { nd = __ldg(a.dst + e0 + lt); ns = __ldg(a.src + e0 + lt); }
        // DIRECT: endpoint ids run two tiles ahead and the row's five input scalars one tile ahead of their use, so that
        // neither load sits on the tile's chain
        int32_t n2d = -1, n2s = 0;
        float2 nxi = make_float2(0.f, 0.f), nxj = nxi;
        float nee = 0.f;
        if (DIRECT) {
            if (e0 + TILE + lt < e1) { n2d = __ldg(a.dst + e0 + TILE + lt); n2s = __ldg(a.src + e0 + TILE + lt); }
            if (nd >= 0) {
                nxi = __ldg(reinterpret_cast<const float2*>(a.x) + nd);
                nxj = __ldg(reinterpret_cast<const float2*>(a.x) + ns);
                nee = __ldg(a.e_prev + e0 + lt);
            }
        }
        for (int32_t t0 = e0; t0 < e1; t0 += TILE) {
            const int nrows = min(TILE, e1 - t0);
            FW_TR(0);
            dst_s[lt] = nd;
            if (!DIRECT) src_s[lt] = ns;
#ifdef GCRL_FWD_L2_PREFETCH
            // Experiment, NOT kept: next tile's e_{l-1} rows -> L2 at the top of this tile, so that the TMA load issued at its
            // end lands from L2.  Measured 3.50 -> 3.59 ms per launch (and a lower power-capped clock): the landing wait is
            // 0.5-0.8 k of a 25 k-cycle chain, and the extra requests compete with the other three groups' streams.
            if (!FIRST && lt == 96 && t0 + TILE < e1)
                l2_prefetch_bulk(a.e_prev + (int64_t)(t0 + TILE) * 64, (uint32_t)min(TILE, e1 - t0 - TILE) * 256u);
#endif
            const bool row_ok = nd >= 0;               // DIRECT: this thread's row lies inside the item
            const float2 xi = nxi, xj = nxj;
            const float ee = nee;
            named_sync(bar_id, 128);
            {
                const bool st = lt > 0 && lt < nrows && dst_s[lt] != dst_s[lt - 1];
                const uint32_t bits = __ballot_sync(0xffffffffu, st);
                if (lane == 0) S.start_bits[wg][w] = bits;
                if (DIRECT) {
                    nd = n2d; ns = n2s;
                    n2d = -1; n2s = 0;
                    if (t0 + 2 * TILE + lt < e1) { n2d = __ldg(a.dst + t0 + 2 * TILE + lt); n2s = __ldg(a.src + t0 + 2 * TILE + lt); }
                    if (nd >= 0) {
                        nxi = __ldg(reinterpret_cast<const float2*>(a.x) + nd);
                        nxj = __ldg(reinterpret_cast<const float2*>(a.x) + ns);
                        nee = __ldg(a.e_prev + t0 + TILE + lt);
                    }
                } else {
                nd = -1; ns = 0;
                if (t0 + TILE + lt < e1) { nd = __ldg(a.dst + t0 + TILE + lt); ns = __ldg(a.src + t0 + TILE + lt); }
                }
            }
            if (DIRECT) {
                // the previous tile's carry chain, under the prefetches just issued
                if (chain_pending) {
                    if (lt < 64) tile_reduce_chain_t(prev_nrows, red, carry, a.out, lt);
                    chain_pending = false;
                }
                // ---- hid = relu((A x_i + b1) + B x_j + C e) -> operand tile (thread = row); same operation order as
                // k_node_proj_first (Pi, Pj) followed by the gather path ((Pi + Pj) + C e) ----------------------------
                if (ATM) {
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        float y[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float4 wa = wA[32 * half + j];
                            const float2 wc = wC[32 * half + j];
                            const float pi = fmaf(wa.y, xi.y, fmaf(wa.x, xi.x, wc.y));
                            const float pj = fmaf(wa.w, xj.y, fmaf(wa.z, xj.x, 0.f));
                            const float z = fmaf(ee, wc.x, __fadd_rn(pi, pj));
                            y[j] = row_ok ? fmaxf(z, 0.f) : 0.f;
                        }
                        tmem_st_split32<X3>(t_acc1 + lane_addr + 32 * half, y);
                    }
                    tmem_st_wait();
                } else {
#pragma unroll 1
                for (int cb = 0; cb < 8; ++cb) {
                    float y[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 wa = wA[8 * cb + j];
                        const float2 wc = wC[8 * cb + j];
                        const float pi = fmaf(wa.y, xi.y, fmaf(wa.x, xi.x, wc.y));
                        const float pj = fmaf(wa.w, xj.y, fmaf(wa.z, xj.x, 0.f));
                        const float z = fmaf(ee, wc.x, __fadd_rn(pi, pj));
                        y[j] = row_ok ? fmaxf(z, 0.f) : 0.f;
                    }
                    store_split8<X3>(y, Bh, Bl, lt, cb);
                }
                }
                FW_TR(1);
            } else {
            // ---- Pi[dst] + Pj[src] -> acc1 (TMEM), two 16-lane fragment groups per warp ------------------------
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
                const int ra = 32 * w + 16 * g + (lane >> 2), rb = ra + 8;
                const int32_t da = dst_s[ra], db = dst_s[rb];
                const bool va = da >= 0, vb = db >= 0;
                const float2* pja = reinterpret_cast<const float2*>(a.Pj + (int64_t)src_s[ra] * 64 + cp);
                const float2* pjb = reinterpret_cast<const float2*>(a.Pj + (int64_t)src_s[rb] * 64 + cp);
                float2 ya[8], yb[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) { ya[c] = __ldg(pja + 4 * c); yb[c] = __ldg(pjb + 4 * c); }
#ifndef GCRL_FWD_CHAIN_AT_END
                // the previous tile's carry chain (64 threads: a few dozen dependent shared-memory reads, the occasional
                // aggregate write) runs here, under the first fragment group's Pj rows in flight, instead of at the end of
                // its own tile where the other two warps idled at the next top-of-tile barrier
                if (g == 0 && chain_pending) {
                    if (lt < 64) tile_reduce_chain_t(prev_nrows, red, carry, a.out, lt);
                    chain_pending = false;
                }
#endif
                if (va && da != cur_d) {          // new segment: refresh the cached Pi row
                    cur_d = da;
                    const float2* pia = reinterpret_cast<const float2*>(a.Pi + (int64_t)da * 64 + cp);
#pragma unroll
                    for (int c = 0; c < 8; ++c) piv[c] = __ldg(pia + 4 * c);
                }
                const bool b_other = vb && db != cur_d;
                const float2* pib = reinterpret_cast<const float2*>(a.Pi + (int64_t)(vb ? db : 0) * 64 + cp);
                float v[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float2 xb = piv[c];
                    if (b_other) xb = __ldg(pib + 4 * c);
                    v[4 * c + 0] = va ? piv[c].x + ya[c].x : 0.f;
                    v[4 * c + 1] = va ? piv[c].y + ya[c].y : 0.f;
                    v[4 * c + 2] = vb ? xb.x + yb[c].x : 0.f;
                    v[4 * c + 3] = vb ? xb.y + yb[c].y : 0.f;
                }
                if (FIRST) {
                    // block 1: C e is De FMAs per element, folded in here ((Pi + Pj) + C e)
                    for (int d = 0; d < a.De; ++d) {
                        const float ea = va ? a.e_prev[(int64_t)(t0 + ra) * a.De + d] : 0.f;
                        const float eb = vb ? a.e_prev[(int64_t)(t0 + rb) * a.De + d] : 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float2 cc = *reinterpret_cast<const float2*>(&Cf[d][8 * c + cp]);
                            v[4 * c + 0] = fmaf(ea, cc.x, v[4 * c + 0]);
                            v[4 * c + 1] = fmaf(ea, cc.y, v[4 * c + 1]);
                            v[4 * c + 2] = fmaf(eb, cc.x, v[4 * c + 2]);
                            v[4 * c + 3] = fmaf(eb, cc.y, v[4 * c + 3]);
                        }
                    }
                }
                tmem_st_16x256b_x8(t_acc1 + ((uint32_t)(32 * w + 16 * g) << 16), v);
            }
            tmem_st_wait();
            FW_TR(1);
            if (!FIRST) {
                // ---- e_{l-1}: fp32 landing tile -> bf16 hi/lo operand tile, in place through registers ----------
                mbar_wait_bounded(bar_tma, ph_tma);
                ph_tma ^= 1;
                FW_TR(2);
                if (ATM) {
                    // thread = tile row: its 64 values -> bf16 hi | lo columns of the G2 accumulator region (dead until G2)
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        float v[32];
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            *reinterpret_cast<float4*>(&v[4 * c]) = *reinterpret_cast<const float4*>(l_ptr(M, lt, 8 * half + c));
                        tmem_st_split32<X3>(t_acc2 + lane_addr + 32 * half, v);
                    }
                    tmem_st_wait();
                } else {
                float4 x[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int t = k * 128 + lt;
                    x[k] = *reinterpret_cast<const float4*>(l_ptr(M, t >> 4, t & 15));
                }
                // every row has been read before any operand row is written.  (A row's 16 readers -- one half-warp -- are also its
                // writers, so a warp barrier would do, and it does in the backward; here it measured 10 % SLOWER, 3.70 vs 3.34 ms
                // per launch: the group-wide barrier keeps the four warps' conversion bursts from interleaving with other groups')
                named_sync(bar_id, 128);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int t = k * 128 + lt;
                    const int r = t >> 4, c4 = t & 15;
                    uint2 h, l;
                    split4<X3>(x[k], h, l);
                    const uint32_t off = bf_off(r, c4 >> 1) + (c4 & 1) * 8;
                    *reinterpret_cast<uint2*>(T + off) = h;
                    if (X3) *reinterpret_cast<uint2*>(T + 16384 + off) = l;
                }
                }
            }
            if (!ATM) fence_proxy_async();
            tc_fence_before();
            named_sync(bar_id, 128);
            tc_fence_after();
            FW_TR(3);
            if (!FIRST) {
                if (w == 0 && elect_one()) {
                    if (ATM) issue_gemm_feat_ts<X3>(t_acc1, t_acc2, smem_u32(S.Wc[0]), smem_u32(S.Wc[1]), idesc, true);
                    else issue_gemm_feat<X3>(t_acc1, smem_u32(Bh), smem_u32(Bl), smem_u32(S.Wc[0]), smem_u32(S.Wc[1]), idesc, true);
                    umma_commit(bar_mma);
                }
                mbar_wait_bounded(bar_mma, ph_mma);
                ph_mma ^= 1;
                tc_fence_after();
            }
            FW_TR(4);
            // ---- hid = relu(acc1) -> operand tile (thread = row, two 32-column halves) ----------------------------
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                float acc[32];
                tmem_ld32(t_acc1 + lane_addr + 32 * half, acc);
                if (ATM) {
                    // relu in place: the 32 columns just read take the 16 hi + 16 lo columns of the same K range
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = fmaxf(acc[j], 0.f);
                    tmem_st_split32<X3>(t_acc1 + lane_addr + 32 * half, acc);
                } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float y[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[j] = fmaxf(acc[8 * c + j], 0.f);
                    store_split8<X3>(y, Bh, Bl, lt, 4 * half + c);
                }
                }
            }
            if (ATM) tmem_st_wait();
            }   // !DIRECT
            if (!ATM) fence_proxy_async();
            tc_fence_before();
            named_sync(bar_id, 128);
            FW_TR(5);
            if (w == 0 && elect_one()) {
                tc_fence_after();
                if (ATM) issue_gemm_feat_ts<X3>(t_acc2, t_acc1, smem_u32(S.W2[0]), smem_u32(S.W2[1]), idesc, false);
                else
                issue_gemm_feat<X3>(t_acc2, smem_u32(Bh), smem_u32(Bl), smem_u32(S.W2[0]), smem_u32(S.W2[1]), idesc, false);
                umma_commit(bar_mma);
            }
            mbar_wait_bounded(bar_mma, ph_mma);
            ph_mma ^= 1;
            tc_fence_after();
            FW_TR(6);
            // ---- m = acc2 + b2 -> M (fp32 staging over the operand tile: G2 has finished reading it) ---------------
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                float acc[32];
                tmem_ld32(t_acc2 + lane_addr + 32 * half, acc);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 bb = *reinterpret_cast<const float4*>(&S.b2[32 * half + 4 * c]);
                    *reinterpret_cast<float4*>(l_ptr(M, lt, 8 * half + c)) =
                        make_float4(acc[4 * c] + bb.x, acc[4 * c + 1] + bb.y, acc[4 * c + 2] + bb.z, acc[4 * c + 3] + bb.w);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            named_sync(bar_id, 128);
            FW_TR(7);
            // ---- store e_l + segmented PNA reduction -------------------------------------------------------------------
            const bool tma_store = a.m_out && nrows == TILE;
            if (tma_store) {
                if (lt == 0) {
                    tma_store_2d(&tm_out, 0, t0, M);
                    tma_store_2d(&tm_out, 32, t0, M + 128 * 32);
                    bulk_commit();
                }
            } else if (a.m_out) {
                float4* gp = reinterpret_cast<float4*>(a.m_out + (int64_t)t0 * 64);
#pragma unroll 4
                for (int k = 0; k < 16; ++k) {
                    const int idx = k * 128 + lt;
                    const int r = idx >> 4, c4 = idx & 15;
                    if (r < nrows) __stcs(gp + idx, *reinterpret_cast<const float4*>(l_ptr(M, r, c4)));
                }
            }
            tile_reduce_scan_sw_t<ARG>(M, dst_s, S.start_bits[wg], nrows, t0, red, a.out, lt);
            tile_reduce_scan_sw_t<ARG>(M, dst_s, S.start_bits[wg], nrows, t0, red, a.out, lt + 128);
            FW_TR(8);
            if (tma_store && lt == 0) bulk_wait_read0();
            fence_proxy_async();                   // generic reads of M are ordered before the next TMA load into it
            named_sync(bar_id, 128);
            if (!FIRST && lt == 64 && t0 + TILE < e1) {
                mbar_arrive_expect_tx(bar_tma, 128 * 64 * 4);
                tma_load_2d(T, &tm_in, 0, t0 + TILE, bar_tma);
                tma_load_2d(T + 16384, &tm_in, 32, t0 + TILE, bar_tma);
            }
            FW_TR(9);
#ifdef GCRL_FWD_CHAIN_AT_END
            if (lt < 64) tile_reduce_chain_t(nrows, red, carry, a.out, lt);
#else
            chain_pending = true;
            prev_nrows = nrows;
#endif
            FW_TR(10);
#ifdef GCRL_FWD_TRACE_BUILD
            ++tr_tiles;
#endif
        }
        if (lt < 64) {
            if (chain_pending) tile_reduce_chain_t(prev_nrows, red, carry, a.out, lt);
            carry_flush_t(carry, a.out, lt);
        }
        named_sync(bar_id, 128);   // reducer scratch quiescent before the next item
    }
#ifdef GCRL_FWD_TRACE_BUILD
    if (tid == 0 && blockIdx.x == 0 && tr_tiles > 0)
        printf("fwd_wg<%d,%d,%d> grp0 %d tiles, cycles/tile: top %lld ids+gather+st %lld waitTMA %lld convert+sync %lld g1 %lld hid+sync %lld g2 %lld m+sync %lld store+scan %lld storewait+sync %lld chain %lld\n",
               (int)FIRST, (int)X3, (int)ARG, tr_tiles, tr_acc[0] / tr_tiles, tr_acc[1] / tr_tiles, tr_acc[2] / tr_tiles, tr_acc[3] / tr_tiles,
               tr_acc[4] / tr_tiles, tr_acc[5] / tr_tiles, tr_acc[6] / tr_tiles, tr_acc[7] / tr_tiles, tr_acc[8] / tr_tiles, tr_acc[9] / tr_tiles,
               tr_acc[10] / tr_tiles);
#endif
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc<512>(S.tmem_base);
}

template <bool FIRST, bool X3, bool ARG, bool DIRECT = false, bool ATM = false>
static int launch_wg_variant(const CUtensorMap& tin, const CUtensorMap& tout, const FwdWgArgs& a, int grid, cudaStream_t st) {
    const int smem = (int)sizeof(FwdWgSmem) + 1024;
    static bool done = false;
    if (!done) {
        GCRL_CUDA(cudaFuncSetAttribute(k_edge_fwd_wg<FIRST, X3, ARG, DIRECT, ATM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        done = true;
    }
    launch_k(k_edge_fwd_wg<FIRST, X3, ARG, DIRECT, ATM>, dim3(grid), dim3(WG_THREADS), smem, st, tin, tout, a);
    return GCRL_OK;
}

// Block 1 straight from the raw inputs (k_edge_fwd_wg<.., DIRECT>): the reference's dimensions only (two vertex features,
// one edge feature); GCRL_FIRST_DIRECT=0 keeps the projection + gather path (A/B switch, read once per process).
bool first_direct_ok(int Dv, int De, const float* x) {
    static const bool on = !(getenv("GCRL_FIRST_DIRECT") && getenv("GCRL_FIRST_DIRECT")[0] == '0');
    return on && Dv == 2 && De == 1 && x != nullptr && (reinterpret_cast<uintptr_t>(x) & 7u) == 0;
}

// Tensor-core edge forward, four warp groups per SM; same contract as launch_edge_fwd (gn_forward.cu).
int launch_edge_fwd_wg(const BlockW& b, bool first, bool x3, const float* Pi, const float* Pj, const float* e_prev,
                       const int32_t* seg_off, const int32_t* src, const int32_t* dst, int64_t N, int64_t E,
                       float* m_out, float* agg, float* var, int32_t* amin, int32_t* amax, cudaStream_t st,
                       const unsigned char* img_w2, const unsigned char* img_wc, const float* x_first) {
    if (E == 0) return GCRL_OK;
    FwdWgArgs a;
    const bool direct = first && first_direct_ok(b.Dv, b.De, x_first);
    a.x = x_first; a.b1 = b.b1;
    a.img_w2 = img_w2; a.img_wc = img_wc;
    a.Pi = Pi; a.Pj = Pj; a.e_prev = e_prev; a.De = b.De;
    a.W1 = b.W1; a.ldW1 = 2 * b.Dv + b.De; a.coff = 2 * b.Dv;
    a.W2 = b.W2; a.b2 = b.b2;
    a.seg_off = seg_off; a.src = src; a.dst = dst; a.N = (int32_t)N; a.E = (int32_t)E;
    a.m_out = m_out;
    a.out.agg = agg; a.out.var = var; a.out.amin = amin; a.out.amax = amax; a.out.seg_off = seg_off;
    CUtensorMap tin, tout;
    int rc = make_tmap_rows64(&tin, first ? Pi : e_prev, first ? N : E, 128);   // FIRST never dereferences tin
    if (rc) return rc;
    rc = make_tmap_rows64(&tout, m_out ? m_out : Pi, m_out ? E : N, 128);
    if (rc) return rc;
    // Work items are runs of whole segments of about item_edges edges; one warp group walks an item tile by tile, items are
    // dealt round-robin.  Round 1 used 2 048-edge items whatever E: a rollout shard of 1 250 x 50-vertex graphs (3.06 M edges)
    // is 2.5 such items per group, so a third of the groups walked 3 items and the rest 2 -- the 0.82 strong-scaling
    // efficiency of the rollouts at N = 8.  Now every group gets the same number k of items:
    //   * inference (no arg tracking): k = 1, ONE contiguous run of E / groups edges per group, like the backward's contiguous
    //     tile ranges (measured on 10 000 x 50-vertex rollouts: 0.447 -> 0.469 of the HBM peak per launch);
    //   * training (stash + arg tracking): the smallest k that keeps an item within ITEM_EDGES; contiguous runs measured
    //     SLOWER here (3.46 -> 3.92 ms per 512 x 200-vertex launch), so the compact streaming window is kept.
    // Small batches (a 64-graph replay minibatch is 70 k edges) end up with one tile per group either way.
    {
        const int64_t groups = (int64_t)sm_count() * WG_N;
        const int64_t k = amin != nullptr ? (E + groups * ITEM_EDGES - 1) / (groups * ITEM_EDGES) : 1;
        int64_t ie = (E + groups * k - 1) / (groups * k);
        if (k == 1) ie = (ie + TILE - 1) / TILE * TILE;
        a.item_edges = (int)(ie < TILE ? TILE : ie);
    }
    const int64_t items = (E + a.item_edges - 1) / a.item_edges;
    int64_t grid = sm_count();
    if (grid > (items + WG_N - 1) / WG_N) grid = (items + WG_N - 1) / WG_N;
    const int tok = prof_begin(first ? GCRL_PROF_EDGE_FWD_FIRST : (m_out ? GCRL_PROF_EDGE_FWD : GCRL_PROF_EDGE_FWD_LAST), st);
    const bool want_arg = amin != nullptr;
    // A operands in tensor memory (k_edge_fwd_wg<.., ATM>); GCRL_FWD_ATM=0: operand tiles in shared memory (A/B switch)
    static const bool atm = !(getenv("GCRL_FWD_ATM") && getenv("GCRL_FWD_ATM")[0] == '0');
#define WG_LAUNCH(F, X, W) rc = atm ? launch_wg_variant<F, X, W, false, true>(tin, tout, a, (int)grid, st) \
                                    : launch_wg_variant<F, X, W>(tin, tout, a, (int)grid, st)
#define WG_LAUNCH_D(X, W) rc = atm ? launch_wg_variant<true, X, W, true, true>(tin, tout, a, (int)grid, st) \
                                   : launch_wg_variant<true, X, W, true>(tin, tout, a, (int)grid, st)
    if (direct) {
        if (x3) { if (want_arg) WG_LAUNCH_D(true, true); else WG_LAUNCH_D(true, false); }
        else    { if (want_arg) WG_LAUNCH_D(false, true); else WG_LAUNCH_D(false, false); }
    } else
    if (first) { if (x3) { if (want_arg) WG_LAUNCH(true, true, true); else WG_LAUNCH(true, true, false); }
                 else    { if (want_arg) WG_LAUNCH(true, false, true); else WG_LAUNCH(true, false, false); } }
    else       { if (x3) { if (want_arg) WG_LAUNCH(false, true, true); else WG_LAUNCH(false, true, false); }
                 else    { if (want_arg) WG_LAUNCH(false, false, true); else WG_LAUNCH(false, false, false); } }
#undef WG_LAUNCH
#undef WG_LAUNCH_D
    prof_end(tok, st);
    if (rc) return rc;
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

}  // namespace gcrl
