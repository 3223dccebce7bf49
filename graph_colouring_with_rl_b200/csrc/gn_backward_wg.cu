// gn_backward_wg.cu -- fused edge kernel of the GN backward pass, two independent tile chains per SM.
//
// Reference: loss.backward() through architecture.py:51-66 (dqn_agent.py:219).  Round 1 went through two earlier
// designs (removed; git history): a single-group kernel and a warp-specialised pipeline whose stages hand ONE tile chain
//   ep1 (hid) -> G3/G4 -> ep3 (dz1) -> G5/G6
// from role to role; shared memory (224 KB) leaves no room to double-buffer hid/dz1 and dm, so consecutive tiles'
// chains cannot overlap and the kernel's time is the chain's latency (5.4 ms of its 8.0 ms per launch with every
// memory side effect switched off; dropping 32 of its 84 UMMAs per tile changes the time by 1 %).
// The forward showed the cheaper cure for a latency-bound chain (gn_forward_wg.cu): more independent chains per SM
// instead of more stages per chain.  Here: TWO 256-thread groups per CTA (one CTA per SM), each walking its own
// contiguous tile range with the phases in program order and named barriers; the 32 KB of weight tiles are shared.
// A group owns three 32 KB tile buffers, each reused along the tile:
//   B1: TMA landing of e_{l-1} (fp32) -> bf16 hi|lo operand in place (G1, G5) -> fp32 staging of d e_{l-1} -> TMA store
//   B2: TMA landing of d e_l          -> hid operand (G3)                   -> dz1 operand (G5, G6)
//   B3: TMA landing of e_l            -> dm operand (G3, G4); dm lives in registers while B1 is converted
//                                     -> fp32 copy of dz1, source of the dPj reduces
// and one 64-column TMEM accumulator that is in turn acc1 (Pi+Pj preloaded by tcgen05.st, + G1), acc4 (G4) and
// acc6 (G6), plus the two weight-gradient accumulators kept for the whole kernel: 192 columns per group.
//   G1  acc += e_{l-1} . C^T  (on top of Pi[dst]+Pj[src])     G4  acc  = dm . W2
//   G3  accW2 += dm^T . hid                                    G5  accC += dz1^T . e_{l-1}      G6  acc = dz1 . C
// The next tile's d e_l / e_l / e_{l-1} loads are issued as soon as G5/G6, the dPj reduces and the store have released
// their buffers.  While G5/G6 run: dPj[src] += dz1 goes out as TMA tensor reduces (cp.reduce.async.bulk.tensor, one-row
// boxes, eight-row boxes for runs of consecutive sources -- LSU reds cost 0.9 ms per launch, these 0.05 ms), dPi[dst] as
// warp transpose-reduced column sums.  Every endpoint id a thread needs is fetched in one batch at the top of the tile:
// next to 225 KB of shared memory the L1 keeps nothing, and a load inside a phase costs that phase an L2 round trip.
// (tight three-instruction mbarrier waits: 5.11 vs 5.32 ms per launch with the counted waits of -DGCRL_MBAR_TIMEOUT, which remain the debugging build)
#include "gn_common.cuh"
#include "ws.cuh"
#include <stdlib.h>

namespace gcrl {

using namespace tc;

bool first_direct_ok(int Dv, int De, const float* x);      // gn_forward_wg.cu

// phase timeline of group 0 of CTA 0 (diagnostic build: GCRL_NVCC_FLAGS=-DGCRL_BWD_TRACE_BUILD): cycles per phase
// averaged over the group's tiles, printed at the end of the kernel
#ifdef GCRL_BWD_TRACE_BUILD
#define BW_TR(k) do { if (tid == 0 && blockIdx.x == 0) { const long long c_ = clock64(); tr_acc[k] += c_ - tr_last; tr_last = c_; } } while (0)
#else
#define BW_TR(k) do { } while (0)
#endif

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, int c0, int c1, const void* smem_src) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(m), "r"(c0), "r"(c1), "r"(smem_u32(smem_src)) : "memory");
}

// ca = g_mean/deg - [var>0] g_std * mean / (deg * std),  cb = [var>0] g_std / (deg * std)
// With de_scatter != null the min / max gradient terms are added to d e_l here (two scattered reds per
// (vertex, column): 1/(n-1) of the elements) so that the edge kernel's dm pass is purely dense:
//   d e_l[argmin[v,c], c] += g_min[v,c],  d e_l[argmax[v,c], c] += g_max[v,c].
__global__ void k_pna_bwd_coef(const float* __restrict__ dagg, const float* __restrict__ agg,
                               const float* __restrict__ var, const int32_t* __restrict__ seg_off, int64_t N,
                               float* __restrict__ ca, float* __restrict__ cb, const int32_t* __restrict__ amin,
                               const int32_t* __restrict__ amax, float* __restrict__ de_scatter,
                               float* __restrict__ gb2_vertex, float4* __restrict__ zero_ptr, int64_t zero_n4) {
    pdl_prologue();
    // dPi | dPj of the edge kernel that follows start from zero: cleared here instead of by a memset node (which would
    // also break the programmatic launch chain around it)
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < zero_n4; t += (int64_t)gridDim.x * blockDim.x)
        zero_ptr[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    // gb2_vertex != null (block 5, pipelined kernel): db2 = sum over edges of dm = sum_v (g_mean + g_min + g_max)
    // (the std term sums to zero over a segment), accumulated here instead of in the edge kernel
    float bacc = 0.f;
    const int64_t total = N * 64;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = t >> 6;
        const int c = (int)(t & 63);
        const float deg = (float)(seg_off[v + 1] - seg_off[v]);
        float a = 0.f, b = 0.f;
        if (deg > 0.f) {
            a = dagg[v * 256 + c] / deg;
            if (var[t] > 0.f) {
                b = dagg[v * 256 + 192 + c] / (deg * agg[v * 256 + 192 + c]);
                a -= b * agg[v * 256 + c];
            }
        }
        ca[t] = a;
        cb[t] = b;
        if (gb2_vertex && deg > 0.f) bacc += dagg[v * 256 + c] + dagg[v * 256 + 64 + c] + dagg[v * 256 + 128 + c];
        if (de_scatter && deg > 0.f) {
            atomicAdd(de_scatter + (int64_t)amin[t] * 64 + c, dagg[v * 256 + 64 + c]);
            atomicAdd(de_scatter + (int64_t)amax[t] * 64 + c, dagg[v * 256 + 128 + c]);
        }
    }
    if (gb2_vertex) {       // blockDim = 256 and a grid stride that is a multiple of 64: a thread's column is fixed
        __shared__ float red[256];
        red[threadIdx.x] = bacc;
        __syncthreads();
        if (threadIdx.x < 64)
            atomicAdd(&gb2_vertex[threadIdx.x], red[threadIdx.x] + red[threadIdx.x + 64] + red[threadIdx.x + 128] + red[threadIdx.x + 192]);
    }
}

int launch_pna_bwd_coef(const float* dagg, const float* agg, const float* var, const int32_t* seg_off, int64_t N,
                        float* ca, float* cb, const int32_t* amin, const int32_t* amax, float* de_scatter,
                        float* gb2_vertex, float* zero_ptr, int64_t zero_n, cudaStream_t st) {
    if (N == 0) return GCRL_OK;
    int64_t blocks = (N * 64 + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    launch_k(k_pna_bwd_coef, dim3((int)blocks), dim3(256), 0, st, dagg, agg, var, seg_off, N, ca, cb, amin, amax, de_scatter, gb2_vertex,
             reinterpret_cast<float4*>(zero_ptr), zero_n / 4);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

constexpr int BG_N = 2;                  // groups per CTA
constexpr int BG_T = 256;                // threads per group

struct BwdWgSmem {
    __nv_bfloat16 Wc[2][64 * 64];        // C hi/lo (FIRST: C^T in fp32 [d][o])
    __nv_bfloat16 W2[2][64 * 64];
    unsigned char B1[BG_N][32768];
    unsigned char B2[BG_N][32768];
    unsigned char B3[BG_N][32768];
    float bsum[64];
    uint64_t bar_E[BG_N], bar_M[BG_N], bar_D[BG_N], bar_mma[BG_N];
    uint32_t tmem_base;
};

struct BwdWgArgs {
    const float* Pi; const float* Pj;
    const float* e_prev; int De;
    const float* W1; int ldW1, coff;
    const float* W2;
    const float* x; const float* b1;          // DIRECT (block 1 with 2 vertex features + 1 edge feature): raw inputs
    const int32_t* src; const int32_t* dst;
    int32_t N, E;
    const float* ca; const float* cb;
    const float* m; const float* de;          // e_l and d e_l (the TMA maps' bases; here for the L2 prefetches)
    const float* dagg;
    const int32_t* amin; const int32_t* amax;
    float* dPi; float* dPj;
    float* gW1; float* gW2; float* gb2;
    const unsigned char* img_w2; const unsigned char* img_wc;   // pre-converted weight tiles (weight_image.cu) or null
    int dbg;                                  // timing experiments only (GCRL_WS_DBG): 16 no dPj reds, 32 no dPi reds, 2 no gather
};

// FIRST: block 1 (raw edge features, no d e_0).  NODE: block 5 (no d e_5: dm = ca + cb e_5 + arg terms).
// DIRECT (FIRST only; Dv = 2, De = 1): hid and its relu mask are re-evaluated from the row's five input scalars in
// epilogue 1, in the forward's operation order (gn_forward_wg.cu), instead of gathering Pi[dst] + Pj[src] into TMEM.
template <bool FIRST, bool NODE, bool X3, bool DIRECT = false>
__global__ void __launch_bounds__(BG_N * BG_T, 1) k_edge_bwd_wg(const __grid_constant__ CUtensorMap tm_e,
                                                               const __grid_constant__ CUtensorMap tm_m,
                                                               const __grid_constant__ CUtensorMap tm_de,
                                                               const __grid_constant__ CUtensorMap tm_out,
                                                               const __grid_constant__ CUtensorMap tm_dpj,
                                                               const __grid_constant__ CUtensorMap tm_dpj8,
                                                               const __grid_constant__ CUtensorMap tm_dpi,
                                                               const BwdWgArgs a) {
    pdl_prologue();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    BwdWgSmem& S = *reinterpret_cast<BwdWgSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, grp = tid >> 8, gt = tid & 255, gw = gt >> 5, lane = tid & 31;
    const int sp = gw & 3, half = gw >> 2;          // TMEM sub-partition (rows 32sp..) and column half
    const int row = 32 * sp + lane;                 // this thread's tile row in the epilogues
    const int bar_id = 1 + grp;

    if (tid == 0) {
        for (int g = 0; g < BG_N; ++g) {
            mbar_init(&S.bar_E[g], 1); mbar_init(&S.bar_M[g], 1); mbar_init(&S.bar_D[g], 1); mbar_init(&S.bar_mma[g], 1);
        }
        fence_barrier_init();
        tma_prefetch_desc(&tm_m);
        if (!FIRST) { tma_prefetch_desc(&tm_e); tma_prefetch_desc(&tm_out); }
        if (!NODE) tma_prefetch_desc(&tm_de);
    }
    if (tid < 32) tmem_alloc<512>(&S.tmem_base);
    if (a.img_w2) copy_image(S.W2, a.img_w2, 16384, tid, BG_N * BG_T);
    else load_weight_tile(a.W2, 64, 0, S.W2[0], S.W2[1], tid, BG_N * BG_T);
    if (!FIRST) {
        if (a.img_wc) copy_image(S.Wc, a.img_wc, 16384, tid, BG_N * BG_T);
        else load_weight_tile(a.W1, a.ldW1, a.coff, S.Wc[0], S.Wc[1], tid, BG_N * BG_T);
    }
    float (*Cf)[64] = reinterpret_cast<float (*)[64]>(&S.Wc[0][0]);   // FIRST: C^T fp32 [d][o]
    // DIRECT: per output column o, (A[o][0], A[o][1], B[o][0], B[o][1]) and (C[o][0], b1[o])
    float4* const wA = reinterpret_cast<float4*>(&S.Wc[0][0]);
    float2* const wC = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(&S.Wc[0][0]) + 1024);
    if (DIRECT && tid < 64) {
        const float* w = a.W1 + tid * a.ldW1;
        wA[tid] = make_float4(w[0], w[1], w[2], w[3]);
        wC[tid] = make_float2(w[4], a.b1[tid]);
    }
    if (FIRST) {
        if (!DIRECT)
        for (int t = tid; t < a.De * 64; t += BG_N * BG_T) Cf[t >> 6][t & 63] = a.W1[(t & 63) * a.ldW1 + a.coff + (t >> 6)];
        for (int t = tid; t < BG_N * 32768 / 16; t += BG_N * BG_T)   // zero-padded raw edge-feature operand tiles
            reinterpret_cast<uint4*>(&S.B1[0][0])[t] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid < 64) S.bsum[tid] = 0.f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t t_acc = S.tmem_base + 192 * grp, t_w2 = t_acc + 64, t_c = t_acc + 128;
    const uint32_t lane_addr = (uint32_t)(32 * sp) << 16;
    const uint32_t id_feat = make_idesc_bf16(128, 64, false, false);
    const uint32_t id_featT = make_idesc_bf16(128, 64, false, true);
    const uint32_t id_rows = make_idesc_bf16(64, 64, true, true);
    const uint32_t wc_hi = smem_u32(S.Wc[0]), wc_lo = smem_u32(S.Wc[1]);
    const uint32_t w2_hi = smem_u32(S.W2[0]), w2_lo = smem_u32(S.W2[1]);
    unsigned char* const B1 = S.B1[grp];
    unsigned char* const B2 = S.B2[grp];
    unsigned char* const B3 = S.B3[grp];
    const uint32_t b1_hi = smem_u32(B1), b1_lo = b1_hi + 16384;
    const uint32_t b2_hi = smem_u32(B2), b2_lo = b2_hi + 16384;
    const uint32_t b3_hi = smem_u32(B3), b3_lo = b3_hi + 16384;
    __nv_bfloat16* const B1h = reinterpret_cast<__nv_bfloat16*>(B1);
    __nv_bfloat16* const B1l = reinterpret_cast<__nv_bfloat16*>(B1 + 16384);
    __nv_bfloat16* const B2h = reinterpret_cast<__nv_bfloat16*>(B2);
    __nv_bfloat16* const B2l = reinterpret_cast<__nv_bfloat16*>(B2 + 16384);
    float* const B1f = reinterpret_cast<float*>(B1);
    const float* const B2f = reinterpret_cast<const float*>(B2);
    const float* const B3f = reinterpret_cast<const float*>(B3);
    uint64_t* const bar_E = &S.bar_E[grp];
    uint64_t* const bar_M = &S.bar_M[grp];
    uint64_t* const bar_D = &S.bar_D[grp];
    uint64_t* const bar_mma = &S.bar_mma[grp];
    uint32_t ph_E = 0, ph_M = 0, ph_D = 0, ph_mma = 0;

    // contiguous tile ranges: a CTA (and each of its groups) stays inside one or two graphs, so CTAs do not fight over
    // the same dPi / dPj rows and the Pi / Pj / coefficient gathers stay in L1 / L2
    const int64_t n_tiles = ((int64_t)a.E + TILE - 1) / TILE;
    const int64_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t cta_first = (int64_t)blockIdx.x * per_cta;
    const int n_cta = (int)max((int64_t)0, min(per_cta, n_tiles - cta_first));
    const int n0 = (n_cta + 1) / 2;
    const int64_t tile_first = cta_first + (grp ? n0 : 0);
    const int n_mine = grp ? n_cta - n0 : n0;
    auto tile_t0 = [&](int i) { return (int32_t)((tile_first + i) * TILE); };

    const uint64_t pol = l2_policy_evict_first();
    auto load_E = [&](int i) {
        mbar_arrive_expect_tx(bar_E, 32768);
        tma_load_2d_hint(B1, &tm_e, 0, tile_t0(i), bar_E, pol);
        tma_load_2d_hint(B1 + 16384, &tm_e, 32, tile_t0(i), bar_E, pol);
    };
    auto load_M = [&](int i) {
        mbar_arrive_expect_tx(bar_M, 32768);
        tma_load_2d_hint(B3, &tm_m, 0, tile_t0(i), bar_M, pol);
        tma_load_2d_hint(B3 + 16384, &tm_m, 32, tile_t0(i), bar_M, pol);
    };
    auto load_D = [&](int i) {
        mbar_arrive_expect_tx(bar_D, 32768);
        tma_load_2d_hint(B2, &tm_de, 0, tile_t0(i), bar_D, pol);
        tma_load_2d_hint(B2 + 16384, &tm_de, 32, tile_t0(i), bar_D, pol);
    };
    if (gt == 64 && n_mine > 0) {
        load_M(0);
        if (!NODE) load_D(0);
        if (!FIRST) load_E(0);
    }

    const int cp = 2 * (lane & 3);
    const int c4 = gt & 15;
    int32_t cur_d = -1;                            // destination whose Pi row is cached in piv
    float2 piv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) piv[c] = make_float2(0.f, 0.f);
    int32_t cur_v = -1;                            // destination whose coefficients are cached in A4 / B4 (...)
    float4 A4 = make_float4(0.f, 0.f, 0.f, 0.f), B4 = A4, Gmn = A4, Gmx = A4;
    int4 Imn = make_int4(-1, -1, -1, -1), Imx = Imn;
    float bs0 = 0.f, bs1 = 0.f, bs2 = 0.f, bs3 = 0.f;

#ifdef GCRL_BWD_TRACE_BUILD
    long long tr_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tr_last = clock64();
#endif
    // endpoint ids of the fragment rows (ra, ra + 8) of the gather, one tile ahead of their use
    int32_t nda = -1, ndb = -1, nsa = 0, nsb = 0;
    auto gather_ids = [&](int i) {
        nda = -1; ndb = -1; nsa = 0; nsb = 0;
        if (i < n_mine) {
            const int32_t t0 = tile_t0(i);
            const int nrows = min(TILE, a.E - t0);
            const int ra = 32 * sp + 16 * half + (lane >> 2), rb = ra + 8;
            if (ra < nrows) { nda = __ldg(a.dst + t0 + ra); nsa = __ldg(a.src + t0 + ra); }
            if (rb < nrows) { ndb = __ldg(a.dst + t0 + rb); nsb = __ldg(a.src + t0 + rb); }
        }
    };
    if (!DIRECT) gather_ids(0);
    // DIRECT: the row's endpoint ids one tile ahead of their use, so that its input scalars can be requested at the top of the tile
    int32_t nmy_s = 0, nmy_d = -1;
    auto row_ids = [&](int i) {
        nmy_s = 0; nmy_d = -1;
        if (i < n_mine) {
            const int32_t t0 = tile_t0(i);
            if (t0 + row < a.E) { nmy_s = __ldg(a.src + t0 + row); nmy_d = __ldg(a.dst + t0 + row); }
        }
    };
    if (DIRECT) row_ids(0);
    for (int i = 0; i < n_mine; ++i) {
        const int32_t t0 = tile_t0(i);
        const int nrows = min(TILE, a.E - t0);
        BW_TR(0);
        // every endpoint id this thread needs during the tile, fetched in one batch: with 225 KB of the SM's 256 KB
        // configured as shared memory the L1 keeps next to nothing, and an id load inside a phase costs that phase an
        // L2 round trip (the dm loop spent 5 000 cycles per tile on eight serialised ones)
        int32_t vv[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = (gt >> 4) + 16 * it;
            vv[it] = r < nrows ? __ldg(a.dst + t0 + r) : -1;
        }
#ifndef GCRL_NO_L2_PREFETCH
        // Block 1 only: next tile's e_1 / d e_1 rows -> L2 now; their TMA loads are issued as the buffers come free later in
        // this tile and then land from L2 (4.65 -> 4.23 ms per launch).  Blocks 2-5 measured SLOWER with the same prefetch
        // (5.58 -> 5.83 ms, block 5 6.48 -> 6.87 ms, and a lower power-capped clock), so they do not issue it.
        if (FIRST && i + 1 < n_mine) {
            const int32_t tn = tile_t0(i + 1);
            const uint32_t nb = (uint32_t)min(TILE, a.E - tn) * 256u;
            if (gt == 32) l2_prefetch_bulk(a.m + (int64_t)tn * 64, nb);
            if (!NODE && gt == 160) l2_prefetch_bulk(a.de + (int64_t)tn * 64, nb);
            if (!FIRST && gt == 224) l2_prefetch_bulk(a.e_prev + (int64_t)tn * 64, nb);
        }
#endif
        const bool valid = row < nrows;
        int32_t my_src, my_dst;
        float2 xi = make_float2(0.f, 0.f), xj = xi;
        float ee = 0.f;
        if (DIRECT) {
            my_src = nmy_s; my_dst = nmy_d;
            row_ids(i + 1);
            if (valid) {
                xi = __ldg(reinterpret_cast<const float2*>(a.x) + my_dst);
                xj = __ldg(reinterpret_cast<const float2*>(a.x) + my_src);
                ee = __ldg(a.e_prev + t0 + row);
            }
        } else {
            my_src = valid ? __ldg(a.src + t0 + row) : 0;
            my_dst = valid ? __ldg(a.dst + t0 + row) : -1;
        }
        // ---- Pi[dst] + Pj[src] (+ C e in block 1) -> acc (TMEM), 16x256b fragment layout ------------------------------
        if (!DIRECT) {
            const int ra = 32 * sp + 16 * half + (lane >> 2), rb = ra + 8;
            const bool va = ra < nrows && !(a.dbg & 2), vb = rb < nrows && !(a.dbg & 2);
            const int32_t da = nda, db = ndb, sa = nsa, sb = nsb;       // fetched one tile ahead
            gather_ids(i + 1);
            const float2* pja = reinterpret_cast<const float2*>(a.Pj + (int64_t)sa * 64 + cp);
            const float2* pjb = reinterpret_cast<const float2*>(a.Pj + (int64_t)sb * 64 + cp);
            float2 ya[8], yb[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) { ya[c] = __ldg(pja + 4 * c); yb[c] = __ldg(pjb + 4 * c); }
            if (va && da != cur_d) {
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
            tmem_st_16x256b_x8(t_acc + ((uint32_t)(32 * sp + 16 * half) << 16), v);
            tmem_st_wait();
        }
        BW_TR(1);
        // ---- dm = d e_l + ca[v] + cb[v] e_l (+ arg terms in block 5), kept in registers ---------------------------------
        float4 dm[8];
        mbar_wait(bar_M, ph_M);
        ph_M ^= 1;
        if (!NODE) { mbar_wait(bar_D, ph_D); ph_D ^= 1; }
        BW_TR(2);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = (gt >> 4) + 16 * it;
            const int32_t v = vv[it];
            float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v >= 0) {
                if (v != cur_v) {
                    cur_v = v;
                    A4 = __ldg(reinterpret_cast<const float4*>(a.ca + (int64_t)v * 64) + c4);
                    B4 = __ldg(reinterpret_cast<const float4*>(a.cb + (int64_t)v * 64) + c4);
                    if (NODE) {
                        Imn = __ldg(reinterpret_cast<const int4*>(a.amin + (int64_t)v * 64) + c4);
                        Imx = __ldg(reinterpret_cast<const int4*>(a.amax + (int64_t)v * 64) + c4);
                        Gmn = __ldg(reinterpret_cast<const float4*>(a.dagg + (int64_t)v * 256 + 64) + c4);
                        Gmx = __ldg(reinterpret_cast<const float4*>(a.dagg + (int64_t)v * 256 + 128) + c4);
                    }
                }
                const float4 m4 = *reinterpret_cast<const float4*>(l_ptr(B3f, r, c4));
                float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!NODE) d4 = *reinterpret_cast<const float4*>(l_ptr(B2f, r, c4));
                r4.x = fmaf(B4.x, m4.x, A4.x) + d4.x; r4.y = fmaf(B4.y, m4.y, A4.y) + d4.y;
                r4.z = fmaf(B4.z, m4.z, A4.z) + d4.z; r4.w = fmaf(B4.w, m4.w, A4.w) + d4.w;
                if (NODE) {
                    const int32_t slot_e = t0 + r;
                    if (Imn.x == slot_e) r4.x += Gmn.x;
                    if (Imn.y == slot_e) r4.y += Gmn.y;
                    if (Imn.z == slot_e) r4.z += Gmn.z;
                    if (Imn.w == slot_e) r4.w += Gmn.w;
                    if (Imx.x == slot_e) r4.x += Gmx.x;
                    if (Imx.y == slot_e) r4.y += Gmx.y;
                    if (Imx.z == slot_e) r4.z += Gmx.z;
                    if (Imx.w == slot_e) r4.w += Gmx.w;
                }
            }
            bs0 += r4.x; bs1 += r4.y; bs2 += r4.z; bs3 += r4.w;
            dm[it] = r4;
        }
        // ---- e_{l-1}: fp32 landing tile -> registers; then both operand tiles are written in place ----------------------
        float4 x[8];
        BW_TR(3);
        if (!FIRST) {
            mbar_wait(bar_E, ph_E);
            ph_E ^= 1;
            BW_TR(4);
#pragma unroll
            for (int it = 0; it < 8; ++it) x[it] = *reinterpret_cast<const float4*>(l_ptr(B1f, (gt >> 4) + 16 * it, c4));
        }
#ifdef GCRL_CONVERT_NAMED_SYNC
        named_sync(bar_id, BG_T);          // every landing row has been read before any operand row is written
#else
        // Row r of an operand tile overwrites exactly row r of the landing tile (hi: first slab, lo: second), and the 16 threads
        // that read a row (one half-warp) are the ones that write it: a warp barrier replaces the group-wide one (-1 %)
        __syncwarp();
#endif
        BW_TR(5);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = (gt >> 4) + 16 * it;
            const uint32_t off = bf_off(r, c4 >> 1) + (c4 & 1) * 8;
            uint2 h, l;
            if (!FIRST) {
                split4<X3>(x[it], h, l);
                *reinterpret_cast<uint2*>(B1 + off) = h;
                if (X3) *reinterpret_cast<uint2*>(B1 + 16384 + off) = l;
            }
            split4<X3>(dm[it], h, l);
            *reinterpret_cast<uint2*>(B3 + off) = h;
            if (X3) *reinterpret_cast<uint2*>(B3 + 16384 + off) = l;
        }
        if (FIRST && gt < TILE) {
            // zero-padded raw edge features: chunk 0 of the row holds e[0..De)
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = (gt < nrows && j < a.De) ? a.e_prev[(int64_t)(t0 + gt) * a.De + j] : 0.f;
            store_split8<X3>(y, B1h, B1l, gt, 0);
        }
        fence_proxy_async();
        tc_fence_before();
        named_sync(bar_id, BG_T);
        tc_fence_after();
        BW_TR(6);
        if (!FIRST) {
            if (gw == 0 && elect_one()) {
                issue_gemm_feat<X3, true>(t_acc, b1_hi, b1_lo, wc_hi, wc_lo, id_feat, true);       // G1
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, ph_mma);
            ph_mma ^= 1;
            tc_fence_after();
        }
        BW_TR(7);
        // ---- epilogue 1: hid = relu(acc1) -> B2, relu mask in a register ---------------------------------------------------
        uint32_t hm = 0u;
        if (DIRECT) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float y[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 wa = wA[32 * half + 8 * c + j];
                    const float2 wc = wC[32 * half + 8 * c + j];
                    const float pi = fmaf(wa.y, xi.y, fmaf(wa.x, xi.x, wc.y));
                    const float pj = fmaf(wa.w, xj.y, fmaf(wa.z, xj.x, 0.f));
                    const float z = valid ? fmaf(ee, wc.x, __fadd_rn(pi, pj)) : 0.f;
                    y[j] = fmaxf(z, 0.f);
                    hm |= (z > 0.f ? 1u : 0u) << (8 * c + j);
                }
                store_split8<X3>(y, B2h, B2l, row, 4 * half + c);
            }
        } else {
            float acc[32];
            tmem_ld32(t_acc + lane_addr + 32 * half, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float y[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    y[j] = fmaxf(acc[8 * c + j], 0.f);
                    hm |= (acc[8 * c + j] > 0.f ? 1u : 0u) << (8 * c + j);
                }
                store_split8<X3>(y, B2h, B2l, row, 4 * half + c);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        named_sync(bar_id, BG_T);
        BW_TR(8);
        if (gw == 0 && elect_one()) {
            tc_fence_after();
            issue_gemm_rows<X3, true>(t_w2, b3_hi, b3_lo, b2_hi, b2_lo, id_rows, 128, i > 0);       // G3
            issue_gemm_featT<X3, true>(t_acc, b3_hi, b3_lo, w2_hi, w2_lo, id_featT, false);         // G4
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        BW_TR(9);
        // ---- epilogue 3: dz1 = acc4 * relu' -> B2 (G3 has finished with hid) -----------------------------------------------
        float dz[32];
        tmem_ld32(t_acc + lane_addr + 32 * half, dz);
#pragma unroll
        for (int j = 0; j < 32; ++j) dz[j] = ((hm >> j) & 1u) ? dz[j] : 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) store_split8<X3>(&dz[8 * c], B2h, B2l, row, 4 * half + c);
        {
            // fp32 copy of dz1 in B3 (dm is dead: G3/G4 are done) for the dPj reduces below
            float* B3w = reinterpret_cast<float*>(B3);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(l_ptr(B3w, row, 8 * half + c)) = make_float4(dz[4 * c], dz[4 * c + 1], dz[4 * c + 2], dz[4 * c + 3]);
        }
        fence_proxy_async();
        tc_fence_before();
        tc_fence_before();
        named_sync(bar_id, BG_T);
        BW_TR(10);
        if (gw == 0 && elect_one()) {
            tc_fence_after();
            issue_gemm_rows<X3, true>(t_c, b2_hi, b2_lo, b1_hi, b1_lo, id_rows, 128, i > 0);        // G5
            if (!FIRST) issue_gemm_featT<X3, true>(t_acc, b2_hi, b2_lo, wc_hi, wc_lo, id_featT, false);   // G6
            umma_commit(bar_mma);
        }
        // ---- while G5/G6 run: dPj[src] += dz1 (vector reds), dPi[dst] += segment sums (warp transpose-reduce) ---------------
        {
            // dPj[src] += dz1: one TMA tensor reduce per (row, 32-column half) from the fp32 staging tile in B3 (box
            // {32, 1} of the dPj tensor map; SWIZZLE_128B is a function of the shared-memory address, so a one-row box
            // at the row's address undoes the staging tile's XOR pattern).  LSU reds -- 16 or 32 bytes per lane into
            // 32 / 8 different lines per instruction -- cost 0.9 ms per launch; these run in the TMA / L2 reduce path
            // Eight tile rows whose sources are consecutive vertices (the rule inside a segment of a complete graph, except
            // across the gap at the destination itself) go as ONE eight-row box.
            if (!(a.dbg & 16)) {
                const int32_t s_next = __shfl_down_sync(0xffffffffu, my_src, 1);
                const uint32_t okb = __ballot_sync(0xffffffffu, valid && ((lane & 7) == 7 || s_next == my_src + 1));
                const bool run8 = ((okb >> (lane & 24)) & 0xffu) == 0xffu;
                if (run8) {
                    if ((lane & 7) == 0) {
                        tma_reduce_add_2d(&tm_dpj8, 32 * half, my_src, B3 + half * 16384 + row * 128);
                        bulk_commit();
                    }
                } else if (valid) {
                    tma_reduce_add_2d(&tm_dpj, 32 * half, my_src, B3 + half * 16384 + row * 128);
                    bulk_commit();
                }
            }
            const int32_t dprev = __shfl_up_sync(0xffffffffu, my_dst, 1);
            const uint32_t sbits = __ballot_sync(0xffffffffu, valid && lane > 0 && my_dst != dprev);
            const int nb = __popc(sbits);
            const int32_t d0 = __shfl_sync(0xffffffffu, my_dst, 0);
            const int bpos = nb ? (__ffs(sbits) - 1) : 0;
            const int32_t d1 = __shfl_sync(0xffffffffu, my_dst, bpos);
            if (a.dbg & 32) {
            } else if (nb == 0) {
                if (d0 >= 0) {
                    const float sum = warp_colsum32(dz, lane);
                    atomicAdd(a.dPi + (int64_t)d0 * 64 + 32 * half + lane, sum);
                }
            } else if (nb == 1) {
                float t[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) t[j] = lane < bpos ? dz[j] : 0.f;
                const float s0 = warp_colsum32(t, lane);
                atomicAdd(a.dPi + (int64_t)d0 * 64 + 32 * half + lane, s0);
#pragma unroll
                for (int j = 0; j < 32; ++j) dz[j] = (lane >= bpos && valid) ? dz[j] : 0.f;
                const float s1 = warp_colsum32(dz, lane);
                atomicAdd(a.dPi + (int64_t)d1 * 64 + 32 * half + lane, s1);
            } else if (valid) {
                // three or more segments inside these 32 rows (graphs of a few dozen vertices: the reference's training
                // shapes): the row goes out as a one-row TMA reduce from the fp32 staging tile, like dPj -- eight 16-byte LSU
                // reds per lane here made this phase 6-16 k cycles of a 35 k-cycle tile on 15-50-vertex minibatches
#ifdef GCRL_DPI_LSU        // A/B build: the LSU reds this path used before
                float* pi = a.dPi + (int64_t)my_dst * 64 + 32 * half;
#pragma unroll
                for (int c = 0; c < 8; ++c) red_add_v4(pi + 4 * c, dz[4 * c], dz[4 * c + 1], dz[4 * c + 2], dz[4 * c + 3]);
#else
                tma_reduce_add_2d(&tm_dpi, 32 * half, my_dst, B3 + half * 16384 + row * 128);
                bulk_commit();
#endif
            }
        }
        BW_TR(11);
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        BW_TR(12);
        if (!NODE && gt == 64 && i + 1 < n_mine) load_D(i + 1);     // B2 is free
        // ---- epilogue 4: d e_{l-1} = acc6 -> fp32 staging in B1 (G5 has finished with it) -> TMA store -----------------------
        bulk_wait_read0();                             // this thread's dPj reduce has read its staging row
        if (!FIRST) {
            float acc[32];
            tmem_ld32(t_acc + lane_addr + 32 * half, acc);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(l_ptr(B1f, row, 8 * half + c)) = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
            fence_proxy_async();
            tc_fence_before();
            named_sync(bar_id, BG_T);
            BW_TR(13);
            if (gt == 96 && i + 1 < n_mine) load_M(i + 1);          // B3 (dz1 staging) is free: the next tile's e_l
            if (gt == 64) {
                tma_store_2d_hint(&tm_out, 0, t0, B1f, pol);              // rows past E are clipped by the tensor map
                tma_store_2d_hint(&tm_out, 32, t0, B1f + 128 * 32, pol);
                bulk_commit();
                bulk_wait_read0();
                if (i + 1 < n_mine) load_E(i + 1);
            }
        } else {
            tc_fence_before();
            named_sync(bar_id, BG_T);      // every epilogue read of acc precedes the next tile's tcgen05.st
            if (gt == 96 && i + 1 < n_mine) load_M(i + 1);
        }
    }
#ifdef GCRL_BWD_TRACE_BUILD
    if (tid == 0 && blockIdx.x == 0 && n_mine > 0)
        printf("bwd_wg<%d,%d,%d> grp0 %d tiles, cycles/tile: top %lld gather %lld waitMD %lld dm %lld waitE %lld xread+sync %lld write+sync %lld g1 %lld ep1+sync %lld g3g4 %lld ep3+sync %lld reds %lld g5g6 %lld ep4+sync %lld\n",
               (int)FIRST, (int)NODE, (int)X3, n_mine, tr_acc[0] / n_mine, tr_acc[1] / n_mine, tr_acc[2] / n_mine, tr_acc[3] / n_mine, tr_acc[4] / n_mine,
               tr_acc[5] / n_mine, tr_acc[6] / n_mine, tr_acc[7] / n_mine, tr_acc[8] / n_mine, tr_acc[9] / n_mine, tr_acc[10] / n_mine,
               tr_acc[11] / n_mine, tr_acc[12] / n_mine, tr_acc[13] / n_mine);
#endif
    if (!FIRST && gt == 64) bulk_wait0();
    bulk_wait0();
    // ---- weight gradients: TMEM -> global (one atomic pass per group), bias gradient ---------------------------------------------
    tc_fence_after();
    if (n_mine > 0) {
        const int o = 16 * sp + lane;          // accumulator row of an M = 64 product (lanes 0..15 of each sub-partition)
        float v[32];
        tmem_ld32(t_w2 + lane_addr + 32 * half, v);
        if (lane < 16)
#pragma unroll
            for (int j = 0; j < 32; j += 4) red_add_v4(&a.gW2[o * 64 + 32 * half + j], v[j], v[j + 1], v[j + 2], v[j + 3]);
        tmem_ld32(t_c + lane_addr + 32 * half, v);
        if (lane < 16) {
            if (!FIRST) {
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(&a.gW1[o * a.ldW1 + a.coff + 32 * half + j], v[j]);
            } else if (half == 0) {
                for (int j = 0; j < a.De; ++j) atomicAdd(&a.gW1[o * a.ldW1 + a.coff + j], v[j]);
            }
        }
    }
    atomicAdd(&S.bsum[4 * c4 + 0], bs0); atomicAdd(&S.bsum[4 * c4 + 1], bs1);
    atomicAdd(&S.bsum[4 * c4 + 2], bs2); atomicAdd(&S.bsum[4 * c4 + 3], bs3);
    tc_fence_before();
    __syncthreads();
    if (tid < 64) atomicAdd(&a.gb2[tid], S.bsum[tid]);
    if (tid < 32) tmem_dealloc<512>(S.tmem_base);
}

template <bool FIRST, bool NODE, bool X3, bool DIRECT = false>
static int launch_bwd_wg_variant(const CUtensorMap& te, const CUtensorMap& tm, const CUtensorMap& tde, const CUtensorMap& tout,
                                 const CUtensorMap& tdpj, const CUtensorMap& tdpj8, const CUtensorMap& tdpi, const BwdWgArgs& a, int grid,
                                 cudaStream_t st) {
    const int smem = (int)sizeof(BwdWgSmem) + 1024;
    static bool done = false;
    if (!done) {
        GCRL_CUDA(cudaFuncSetAttribute(k_edge_bwd_wg<FIRST, NODE, X3, DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        done = true;
    }
    launch_k(k_edge_bwd_wg<FIRST, NODE, X3, DIRECT>, dim3(grid), dim3(BG_N * BG_T), smem, st, te, tm, tde, tout, tdpj, tdpj8, tdpi, a);
    return GCRL_OK;
}

// Two-group tensor-core edge backward of one block.  m = e_l must be given (the forward stashes e_5 in the tensor-core
// modes, so block 5 is a regular block without a d e_l term); de == nullptr <=> block 5.
int launch_edge_bwd_wg(bool first, bool x3, const float* Pi, const float* Pj, const float* e_prev, int De,
                       const float* W1, int ldW1, int coff, const float* W2, const float* b2, const int32_t* seg_off,
                       const int32_t* src, const int32_t* dst, int64_t N, int64_t E, const float* m, const float* de,
                       const float* ca, const float* cb, const float* dagg, const int32_t* amin, const int32_t* amax,
                       float* de_prev, float* dPi, float* dPj, float* gW1, float* gW2, float* gb2, cudaStream_t st,
                       const unsigned char* img_w2, const unsigned char* img_wc, const float* x_first, const float* b1,
                       int Dv) {
    (void)seg_off; (void)b2;
    if (E == 0) return GCRL_OK;
    GCRL_CHECK_ARG(m != nullptr);
    const bool no_de = (de == nullptr);
    GCRL_CHECK_ARG(!(first && no_de));
    BwdWgArgs a;
    a.Pi = Pi; a.Pj = Pj; a.e_prev = e_prev; a.De = De;
    a.W1 = W1; a.ldW1 = ldW1; a.coff = coff; a.W2 = W2;
    a.src = src; a.dst = dst; a.N = (int32_t)N; a.E = (int32_t)E;
    a.ca = ca; a.cb = cb; a.dagg = dagg; a.amin = amin; a.amax = amax;
    a.m = m; a.de = de;
    a.dPi = dPi; a.dPj = dPj; a.gW1 = gW1; a.gW2 = gW2; a.gb2 = gb2;
    a.img_w2 = img_w2; a.img_wc = img_wc;
    a.x = x_first; a.b1 = b1;
    const bool direct = first && b1 != nullptr && first_direct_ok(Dv, De, x_first);
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("GCRL_WS_DBG"); dbg = e ? atoi(e) : 0; } a.dbg = dbg; }
    CUtensorMap te, tm, tde, tout, tdpj, tdpj8, tdpi;
    int rc;
    // unused maps point at a valid [N,64] array so that encoding never fails
    if ((rc = make_tmap_rows64(&te, first ? Pi : e_prev, first ? N : E, 128))) return rc;
    if ((rc = make_tmap_rows64(&tm, m, E, 128))) return rc;
    if ((rc = make_tmap_rows64(&tde, no_de ? Pi : de, no_de ? N : E, 128))) return rc;
    if ((rc = make_tmap_rows64(&tout, first ? Pi : de_prev, first ? N : E, 128))) return rc;
    if ((rc = make_tmap_rows64(&tdpj, dPj, N, 1))) return rc;
    if ((rc = make_tmap_rows64(&tdpj8, dPj, N, 8))) return rc;
    if ((rc = make_tmap_rows64(&tdpi, dPi, N, 1))) return rc;
    const int64_t tiles = (E + TILE - 1) / TILE;
    const int64_t per_cta = (tiles + sm_count() - 1) / sm_count();
    const int64_t grid = (tiles + per_cta - 1) / per_cta;        // every CTA owns >= 1 tile of a contiguous range
    const int tok = prof_begin(first ? GCRL_PROF_EDGE_BWD_FIRST : (no_de ? GCRL_PROF_EDGE_BWD_LAST : GCRL_PROF_EDGE_BWD), st);
    if (direct) rc = x3 ? launch_bwd_wg_variant<true, false, true, true>(te, tm, tde, tout, tdpj, tdpj8, tdpi, a, (int)grid, st)
                        : launch_bwd_wg_variant<true, false, false, true>(te, tm, tde, tout, tdpj, tdpj8, tdpi, a, (int)grid, st);
    else if (first) rc = x3 ? launch_bwd_wg_variant<true, false, true>(te, tm, tde, tout, tdpj, tdpj8, tdpi, a, (int)grid, st)
                       : launch_bwd_wg_variant<true, false, false>(te, tm, tde, tout, tdpj, tdpj8, tdpi, a, (int)grid, st);
    else if (no_de) rc = x3 ? launch_bwd_wg_variant<false, true, true>(te, tm, tde, tout, tdpj, tdpj8, tdpi, a, (int)grid, st)
                            : launch_bwd_wg_variant<false, true, false>(te, tm, tde, tout, tdpj, tdpj8, tdpi, a, (int)grid, st);
    else rc = x3 ? launch_bwd_wg_variant<false, false, true>(te, tm, tde, tout, tdpj, tdpj8, tdpi, a, (int)grid, st)
                 : launch_bwd_wg_variant<false, false, false>(te, tm, tde, tout, tdpj, tdpj8, tdpi, a, (int)grid, st);
    prof_end(tok, st);
    if (rc) return rc;
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

}  // namespace gcrl
