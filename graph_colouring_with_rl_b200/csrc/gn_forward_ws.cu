// gn_forward_ws.cu -- warp-specialised, pipelined fused edge kernel of the GN forward pass.
//
// Same math and fusion as k_edge_fwd_tc (gn_forward_tc.cu; reference: GNNBlock.message / aggregate,
// architecture.py:51-66 with the first Linear split into per-vertex projections, PNA aggregation of
// principal_neighbourhood_aggregation.py:41-64 fused in), but the phases of consecutive 128-edge tiles overlap:
// one persistent CTA per SM, 24 warps, every warp group owns a stage and hands tiles on through mbarriers.
//
//   warp 0  lane 0   TMA producer: e_{l-1} boxes (32 rows x 64 fp32) into a ring of landing slots
//   warp 1           tcgen05.mma issuer: G1  ACC1[a] += e_{l-1} . C^T  (on top of Pi + Pj), G2  ACC2[b] = hid . W2^T
//   warp 3  lane 0   TMA store of the message tile to e_l
//   WG_A  warps 4-7   Pi[dst] + Pj[src] gathered in the 16x256b fragment layout -> TMEM (tcgen05.st): the
//                     projections never touch shared memory; block 1 folds C e (edge_dim FMAs) in here
//   WG_C  warps 8-11  epilogue 1: hid = relu(ACC1) -> bf16 hi/lo operand tile B2
//   WG_D  warps 12-15 epilogue 2: m = ACC2 + b2 -> fp32 tile M; then e_{l-1} landing boxes of tile i+2 -> bf16 hi/lo
//                     operand tile B1[s]
//   WG_R  warps 16-23 segmented PNA reduction of M (column scan + carry chain, segreduce.cuh): mean|min|max|std,
//                     arg slots for the backward pass; manual store of the partial last tile
//
// A CTA owns ONE contiguous range of whole destination segments (binary search on seg_off at kernel start): no
// cross-CTA combination, no atomics, and the reducer's carry lives in registers for the whole kernel.
// Shared memory (bf16x3): weights 32 + B1 2x32 + B2 32 + M 32 + landing ring 5x8 + reducer scratch 12 = 212 KB.
// TMEM: ACC1 x3, ACC2 x2 (64 columns each).  HBM traffic per edge is unchanged (blocks 2-4: read 256 B, write 256 B).
#include "gn_common.cuh"
#include "ws.cuh"
#include <stdlib.h>

namespace gcrl {

constexpr int FW_NE = 5;
constexpr int FW_THREADS = 768;

struct FwdWsSmem {
    __nv_bfloat16 Wc[2][64 * 64];        // C hi/lo (FIRST: C^T in fp32 [d][o])
    __nv_bfloat16 W2[2][64 * 64];
    unsigned char B1[2][32768];          // e_{l-1} operand (hi 16 KB | lo 16 KB)
    unsigned char B2[32768];             // hid operand
    float M[128 * 64];                   // message tile, fp32 SW128 slabs (TMA store source, reducer input)
    unsigned char ringE[FW_NE][8192];
    RedScratch red;
    float b2[64];
    uint64_t bar_Efull[FW_NE], bar_Eempty[FW_NE];
    uint64_t bar_Pfull[3], bar_acc1full[3], bar_acc1free[3];
    uint64_t bar_B1full[2], bar_B1free[2], bar_acc2full[2], bar_acc2free[2];
    uint64_t bar_B2full, bar_Mfull[2], bar_Mfree[2];       // M is handed over per 32-column slab
    int32_t head_dst[2][4], tail_dst[2][4];                 // reducer bookkeeping per column-half group
    uint32_t tmem_base;
};

struct FwdWsArgs {
    const float* Pi; const float* Pj;
    const float* e_prev; int De;
    const float* W1; int ldW1, coff;
    const float* W2; const float* b2;
    const int32_t* seg_off; const int32_t* src; const int32_t* dst;
    int32_t N, E;
    float* m_out;
    AggOut out;
    long long* trace;
    int dbg;                 // timing experiments only (GCRL_FW_DBG): 1 no gather loads, 2 no reduction scan, 4 no conversion
};

__device__ __forceinline__ long long clock_pinned() {      // not reordered across memory operations
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory");
    return t;
}
#ifdef GCRL_WS_TRACE_BUILD
#define FW_TR(on, role, tile, ev)                                                           \
    do {                                                                                    \
        if (a.trace && (on) && blockIdx.x == 0 && (tile) < 64)                              \
            a.trace[(((role) * 64 + (tile)) * 8) + (ev)] = clock_pinned();                  \
    } while (0)
#else
#define FW_TR(on, role, tile, ev) do { } while (0)
#endif

template <bool FIRST, bool X3, bool WANT_ARG>
__global__ void __launch_bounds__(FW_THREADS, 1) k_edge_fwd_ws(const __grid_constant__ CUtensorMap tm_in,
                                                              const __grid_constant__ CUtensorMap tm_out,
                                                              const FwdWsArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    FwdWsSmem& S = *reinterpret_cast<FwdWsSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3;
    const int wtid = tid & 127;
    const long long t_start = clock64();
    const bool has_out = a.m_out != nullptr;

    if (tid == 0) {
        for (int i = 0; i < FW_NE; ++i) { mbar_init(&S.bar_Efull[i], 1); mbar_init(&S.bar_Eempty[i], 4); }
        for (int i = 0; i < 3; ++i) { mbar_init(&S.bar_Pfull[i], 4); mbar_init(&S.bar_acc1full[i], 1); mbar_init(&S.bar_acc1free[i], 4); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&S.bar_B1full[i], 4);
            mbar_init(&S.bar_B1free[i], 1);
            mbar_init(&S.bar_acc2full[i], 1);
            mbar_init(&S.bar_acc2free[i], 4);
        }
        mbar_init(&S.bar_B2full, 4);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&S.bar_Mfull[i], 4);
            mbar_init(&S.bar_Mfree[i], has_out ? 5 : 4);   // 4 reducer warps of the slab's group (+ the store thread)
        }
        fence_barrier_init();
    }
    if (warp == 22) tmem_alloc<512>(&S.tmem_base);
    load_weight_tile(a.W2, 64, 0, S.W2[0], S.W2[1], tid, FW_THREADS);
    if (!FIRST) load_weight_tile(a.W1, a.ldW1, a.coff, S.Wc[0], S.Wc[1], tid, FW_THREADS);
    float (*Cf)[64] = reinterpret_cast<float (*)[64]>(&S.Wc[0][0]);
    if (FIRST)
        for (int t = tid; t < a.De * 64; t += FW_THREADS) Cf[t >> 6][t & 63] = a.W1[(t & 63) * a.ldW1 + a.coff + (t >> 6)];
    if (tid < 64) S.b2[tid] = a.b2[tid];
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    const uint32_t t_acc1 = tmem, t_acc2 = tmem + 192;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;

    // this CTA's range of whole segments: [e0, e1)
    const int64_t chunk = ((int64_t)a.E + gridDim.x - 1) / gridDim.x;
    const int32_t v0 = seg_lower_bound(a.seg_off, a.N, (int64_t)blockIdx.x * chunk);
    const int32_t v1 = (blockIdx.x + 1 == gridDim.x) ? a.N : seg_lower_bound(a.seg_off, a.N, (int64_t)(blockIdx.x + 1) * chunk);
    const int32_t e0 = a.seg_off[v0], e1 = a.seg_off[v1];
    const int n_cta = (e1 - e0 + TILE - 1) / TILE;          // may be 0
    auto tile_t0 = [&](int i) { return e0 + i * TILE; };

    // landing boxes jlo..jlo+1 of tile j2 -> bf16 hi/lo operand tile B1[j2 & 1]; WG_C converts boxes 0-1, WG_D boxes 2-3
    auto convert = [&](int j2, int jlo, int jn) {
        if (FIRST || j2 >= n_cta) return;
        const int s = j2 & 1;
        unsigned char* b1 = S.B1[s];
        const int nrows = min(TILE, e1 - tile_t0(j2));
        if (j2 >= 2) mbar_wait(&S.bar_B1free[s], ((j2 - 2) >> 1) & 1);      // G1 two tiles ago has read B1[s]
        for (int j = jlo; j < jlo + jn; ++j) {
            if (32 * j < nrows) {
                const int k = 4 * j2 + j;                                        // every tile before the last one is full
                const int slot = k % FW_NE, u = k / FW_NE;
                mbar_wait(&S.bar_Efull[slot], u & 1);
                const unsigned char* box = S.ringE[slot];
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int r = (wtid >> 4) + 8 * it, c4 = wtid & 15;
                    const float4 x = *reinterpret_cast<const float4*>(box + sw_off<32>(r, c4));
                    uint2 h, l;
                    split4<X3>(x, h, l);
                    const uint32_t off = bf_off(32 * j + r, c4 >> 1) + (c4 & 1) * 8;
                    *reinterpret_cast<uint2*>(b1 + off) = h;
                    if (X3) *reinterpret_cast<uint2*>(b1 + 16384 + off) = l;
                }
                warp_arrive(&S.bar_Eempty[slot], lane);
            } else {
                for (int t = wtid; t < 32 * 8; t += 128) {
                    const uint32_t off = bf_off(32 * j + (t >> 3), t & 7);
                    *reinterpret_cast<uint4*>(b1 + off) = make_uint4(0u, 0u, 0u, 0u);
                    if (X3) *reinterpret_cast<uint4*>(b1 + 16384 + off) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
        fence_proxy_async();
        warp_arrive(&S.bar_B1full[s], lane);
    };

    // Warp ids: the scheduler favours high warp ids, so the light, latency-critical roles sit at the top and the
    // ALU-heavy reducer at the bottom: 0-7 WG_R, 8-11 WG_D, 12-15 WG_C, 16-19 WG_A, 20 TMA, 21 MMA, 22 TMEM alloc, 23 store.
    if (warp == 20) {
        // ================= TMA producer ====================================================================
        if (!FIRST && lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            for (int i = 0; i < n_cta; ++i) {
                const int32_t t0 = tile_t0(i);
                const int nrows = min(TILE, e1 - t0);
                for (int j = 0; 32 * j < nrows; ++j) {
                    const int k = 4 * i + j;
                    const int slot = k % FW_NE, u = k / FW_NE;
                    mbar_wait(&S.bar_Eempty[slot], (u & 1) ^ 1);
                    FW_TR(true, 0, i, j);
                    mbar_arrive_expect_tx(&S.bar_Efull[slot], 8192);
                    tma_load_2d_hint(S.ringE[slot], &tm_in, 0, t0 + 32 * j, &S.bar_Efull[slot], pol);
                    tma_load_2d_hint(S.ringE[slot] + 4096, &tm_in, 32, t0 + 32 * j, &S.bar_Efull[slot], pol);
                }
            }
        }
    } else if (warp == 23) {
        // ================= TMA store of full message tiles =====================================================
        if (has_out && lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            for (int i = 0; i < n_cta; ++i) {
                const int32_t t0 = tile_t0(i);
                for (int h = 0; h < 2; ++h) {
                    mbar_wait(&S.bar_Mfull[h], i & 1);
                    if (e1 - t0 >= TILE) {          // the partial last tile is stored by the reducer groups (no clipping at e1)
                        tma_store_2d_hint(&tm_out, 32 * h, t0, S.M + h * 128 * 32, pol);
                        bulk_commit();
                        bulk_wait_read0();
                    }
                    mbar_arrive(&S.bar_Mfree[h]);
                }
                FW_TR(true, 6, i, 2);
            }
            bulk_wait0();
        }
    } else if (warp == 21) {
        // ================= MMA issuer ============================================================================
        const uint32_t id_feat = make_idesc_bf16(128, 64, false, false);
        const uint32_t wc_hi = smem_u32(S.Wc[0]), wc_lo = smem_u32(S.Wc[1]);
        const uint32_t w2_hi = smem_u32(S.W2[0]), w2_lo = smem_u32(S.W2[1]);
        const uint32_t b2_hi = smem_u32(S.B2), b2_lo = b2_hi + 16384;
        auto g1_ready = [&](int j) {
            uint32_t ok = mbar_test_wait(&S.bar_Pfull[j % 3], (j / 3) & 1) ? 1u : 0u;
            if (!FIRST) ok &= mbar_test_wait(&S.bar_B1full[j & 1], (j >> 1) & 1) ? 1u : 0u;
            return __shfl_sync(0xffffffffu, ok, 0) != 0;
        };
        auto issue_g1 = [&](int j) {
            if (FIRST) return;                     // block 1: ACC1 is complete after the gather (WG_C waits on Pfull)
            const int s = j & 1;
            mbar_wait(&S.bar_Pfull[j % 3], (j / 3) & 1);
            mbar_wait(&S.bar_B1full[s], (j >> 1) & 1);
            tc_fence_after();
            FW_TR(lane == 0, 1, j, 0);
            const uint32_t b1_hi = smem_u32(S.B1[s]);
            if (elect_one()) {
                issue_gemm_feat<X3, true>(t_acc1 + 64 * (j % 3), b1_hi, b1_hi + 16384, wc_hi, wc_lo, id_feat, true);
                umma_commit(&S.bar_acc1full[j % 3]);
                umma_commit(&S.bar_B1free[s]);
            }
            __syncwarp();
        };
        if (n_cta > 0) issue_g1(0);
        for (int i = 0; i < n_cta; ++i) {
            bool g1_next = (i + 1 >= n_cta) || FIRST;
            if (!g1_next && g1_ready(i + 1)) { issue_g1(i + 1); g1_next = true; }
            mbar_wait(&S.bar_B2full, i & 1);
            mbar_wait(&S.bar_acc2free[i & 1], ((i >> 1) & 1) ^ 1);
            tc_fence_after();
            FW_TR(lane == 0, 1, i, 1);
            if (elect_one()) {
                issue_gemm_feat<X3, true>(t_acc2 + 64 * (i & 1), b2_hi, b2_lo, w2_hi, w2_lo, id_feat, false);
                umma_commit(&S.bar_acc2full[i & 1]);
            }
            __syncwarp();
            FW_TR(lane == 0, 1, i, 2);
            if (!g1_next) issue_g1(i + 1);
        }
    } else if (warp >= 16 && warp < 20) {
        // ================= WG_A: Pi[dst] + Pj[src] (+ C e for block 1) -> TMEM ====================================
        const int cp = 2 * (lane & 3);
        int32_t nd[4], ns[4];
        auto load_ids = [&](int i) {
#pragma unroll
            for (int t = 0; t < 4; ++t) { nd[t] = -1; ns[t] = 0; }
            if (i < n_cta) {
                const int32_t t0 = tile_t0(i);
                const int nrows = min(TILE, e1 - t0);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int r = 32 * q + 16 * (t >> 1) + 8 * (t & 1) + (lane >> 2);
                    if (r < nrows) { nd[t] = __ldg(a.dst + t0 + r); ns[t] = __ldg(a.src + t0 + r); }
                }
            }
        };
        load_ids(0);
        int32_t cur_d = -1;
        float2 piv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) piv[c] = make_float2(0.f, 0.f);
        for (int i = 0; i < n_cta; ++i) {
            const int a3 = i % 3;
            const int32_t t0 = tile_t0(i);
            int32_t d4[4], s4[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) { d4[t] = nd[t]; s4[t] = ns[t]; }
            FW_TR(wtid == 0, 3, i, 2);
            load_ids(i + 1);
            FW_TR(wtid == 0, 3, i, 3);
            mbar_wait(&S.bar_acc1free[a3], ((i / 3) & 1) ^ 1);
            tc_fence_after();
            FW_TR(wtid == 0, 3, i, 0);
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
                const int ra = 32 * q + 16 * g + (lane >> 2), rb = ra + 8;
                const int32_t da = g ? d4[2] : d4[0], db = g ? d4[3] : d4[1];
                const bool va = da >= 0 && !(a.dbg & 1), vb = db >= 0 && !(a.dbg & 1);
                const float2* pja = reinterpret_cast<const float2*>(a.Pj + (int64_t)(g ? s4[2] : s4[0]) * 64 + cp);
                const float2* pjb = reinterpret_cast<const float2*>(a.Pj + (int64_t)(g ? s4[3] : s4[1]) * 64 + cp);
                float2 ya[8], yb[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    ya[c] = va ? __ldg(pja + 4 * c) : make_float2(0.f, 0.f);
                    yb[c] = vb ? __ldg(pjb + 4 * c) : make_float2(0.f, 0.f);
                }
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
                    const float2 xb = b_other ? __ldg(pib + 4 * c) : piv[c];
                    v[4 * c + 0] = va ? piv[c].x + ya[c].x : 0.f;
                    v[4 * c + 1] = va ? piv[c].y + ya[c].y : 0.f;
                    v[4 * c + 2] = vb ? xb.x + yb[c].x : 0.f;
                    v[4 * c + 3] = vb ? xb.y + yb[c].y : 0.f;
                }
                if (FIRST) {
                    float ce[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) ce[c] = 0.f;
                    for (int d = 0; d < a.De; ++d) {
                        const float ea = va ? a.e_prev[(int64_t)(t0 + ra) * a.De + d] : 0.f;
                        const float eb = vb ? a.e_prev[(int64_t)(t0 + rb) * a.De + d] : 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float2 cc = *reinterpret_cast<const float2*>(&Cf[d][8 * c + cp]);
                            ce[4 * c + 0] = fmaf(ea, cc.x, ce[4 * c + 0]);
                            ce[4 * c + 1] = fmaf(ea, cc.y, ce[4 * c + 1]);
                            ce[4 * c + 2] = fmaf(eb, cc.x, ce[4 * c + 2]);
                            ce[4 * c + 3] = fmaf(eb, cc.y, ce[4 * c + 3]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] += ce[c];
                }
                tmem_st_16x256b_x8(t_acc1 + 64 * a3 + ((uint32_t)(32 * q + 16 * g) << 16), v);
                FW_TR(wtid == 0, 3, i, 5 + g);
            }
            FW_TR(wtid == 0, 3, i, 4);
            tmem_st_wait();
            tc_fence_before();
            FW_TR(wtid == 0, 3, i, 1);
            warp_arrive(&S.bar_Pfull[a3], lane);
            FW_TR(wtid == 0, 3, i, 7);
        }
    } else if (warp >= 12 && warp < 16) {
        // ================= WG_C: epilogue 1, hid = relu(ACC1) -> B2 ====================================================
        const int row = 32 * q + lane;
        __nv_bfloat16* b2h = reinterpret_cast<__nv_bfloat16*>(S.B2);
        __nv_bfloat16* b2l = reinterpret_cast<__nv_bfloat16*>(S.B2 + 16384);
        for (int i = 0; i < n_cta; ++i) {
            const int a3 = i % 3, pa = (i / 3) & 1;
            if (FIRST) mbar_wait(&S.bar_Pfull[a3], pa);
            else mbar_wait(&S.bar_acc1full[a3], pa);
            tc_fence_after();
            FW_TR(wtid == 0, 5, i, 0);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                float acc[32];
                tmem_ld32(t_acc1 + 64 * a3 + lane_addr + 32 * half, acc);
                if (half == 0 && i > 0) mbar_wait(&S.bar_acc2full[(i - 1) & 1], ((i - 1) >> 1) & 1);   // G2 of the previous tile read B2
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float x[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] = fmaxf(acc[8 * c + j], 0.f);
                    store_split8<X3>(x, b2h, b2l, row, 4 * half + c);
                }
            }
            tc_fence_before();
            warp_arrive(&S.bar_acc1free[a3], lane);
            fence_proxy_async();
            FW_TR(wtid == 0, 5, i, 2);
            warp_arrive(&S.bar_B2full, lane);
        }
    } else if (warp >= 8 && warp < 12) {
        // ================= WG_D: epilogue 2 (m = ACC2 + b2 -> M) and the e_{l-1} operand tiles ==============================
        const int row = 32 * q + lane;
        convert(0, 0, 4);
        convert(1, 0, 4);
        for (int i = 0; i < n_cta; ++i) {
            const int b = i & 1;
            mbar_wait(&S.bar_acc2full[b], (i >> 1) & 1);
            tc_fence_after();
            FW_TR(wtid == 0, 4, i, 2);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                float acc[32];
                tmem_ld32(t_acc2 + 64 * b + lane_addr + 32 * half, acc);
                if (i > 0) mbar_wait(&S.bar_Mfree[half], (i - 1) & 1);      // reducer group and store are done with this slab
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 bb = *reinterpret_cast<const float4*>(&S.b2[32 * half + 4 * c]);
                    *reinterpret_cast<float4*>(l_ptr(S.M, row, 8 * half + c)) =
                        make_float4(acc[4 * c] + bb.x, acc[4 * c + 1] + bb.y, acc[4 * c + 2] + bb.z, acc[4 * c + 3] + bb.w);
                }
                fence_proxy_async();
                if (half == 1) {
                    tc_fence_before();
                    warp_arrive(&S.bar_acc2free[b], lane);
                }
                FW_TR(wtid == 0, 4, i, 3);
                warp_arrive(&S.bar_Mfull[half], lane);
            }
            convert(i + 2, 0, 4);
        }
    } else if (warp < 8) {
        // ================= WG_R0 / WG_R1: segmented PNA reduction, one group of 4 warps per 32-column slab of M ===========
        const int grp = warp >= 4 ? 1 : 0;
        const int lt = tid & 127;                  // thread inside the group
        const int col = 32 * grp + (lt & 31), qtr = lt >> 5;
        int32_t* hd = S.head_dst[grp];
        int32_t* td = S.tail_dst[grp];
        Carry carry;
        carry.dst = -1;
        stat_init(carry.s);
        for (int i = 0; i < n_cta; ++i) {
            const int32_t t0 = tile_t0(i);
            const int nrows = min(TILE, e1 - t0);
            // destination ids / segment-start bits of this thread's quarter (every quarter warp computes its own word)
            const int rr = 32 * qtr + (lt & 31);
            const int32_t d = rr < nrows ? __ldg(a.dst + t0 + rr) : -1;
            const int32_t dp = (rr > 0 && rr < nrows) ? __ldg(a.dst + t0 + rr - 1) : d;
            const uint32_t starts = __ballot_sync(0xffffffffu, rr > 0 && rr < nrows && d != dp) & ~1u;
            const int r0 = 32 * qtr, r1 = min(r0 + 32, nrows);
            named_sync(3 + grp, 128);              // the previous tile's chain is done with the scratch
            mbar_wait(&S.bar_Mfull[grp], i & 1);
            FW_TR(lt == 0 && grp == 0, 6, i, 0);
            if (r0 < r1 && !(a.dbg & 2)) {
                const int c4 = col >> 2, cc = c4 & 7;
                const unsigned char* base = reinterpret_cast<const unsigned char*>(S.M) + (c4 >> 3) * (128 * 128) + (col & 3) * 4;
                const int32_t hd0 = __shfl_sync(0xffffffffu, d, 0);
                const int nstart = __popc(starts);
                if (nstart <= 1) {
                    // at most one segment boundary inside the quarter: rows [r0, ra) go to the head partial A, rows
                    // [ra, r1) to the tail partial B.  Groups of 8 rows that lie wholly on one side run straight-line.
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
                                if (WANT_ARG) stat_add(A, v[k], t0 + rb + k);
                                else stat_add_noarg(A, v[k]);
                            }
                        } else if (rb >= ra && rb + 8 <= r1) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                if (WANT_ARG) stat_add(B, v[k], t0 + rb + k);
                                else stat_add_noarg(B, v[k]);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const int r = rb + k;
                                if (r < ra) {
                                    if (WANT_ARG) stat_add(A, v[k], t0 + r);
                                    else stat_add_noarg(A, v[k]);
                                } else if (r < r1) {
                                    if (WANT_ARG) stat_add(B, v[k], t0 + r);
                                    else stat_add_noarg(B, v[k]);
                                }
                            }
                        }
                    }
                    S.red.f[0][0][qtr][col] = A.sum; S.red.f[0][1][qtr][col] = A.sq;
                    S.red.f[0][2][qtr][col] = A.mn;  S.red.f[0][3][qtr][col] = A.mx;
                    S.red.a[0][0][qtr][col] = A.amn; S.red.a[0][1][qtr][col] = A.amx;
                    const bool has_tail = nstart == 1 && r0 + bl < r1;
                    if (has_tail) {
                        S.red.f[1][0][qtr][col] = B.sum; S.red.f[1][1][qtr][col] = B.sq;
                        S.red.f[1][2][qtr][col] = B.mn;  S.red.f[1][3][qtr][col] = B.mx;
                        S.red.a[1][0][qtr][col] = B.amn; S.red.a[1][1][qtr][col] = B.amx;
                    }
                    const int32_t tdv = __shfl_sync(0xffffffffu, d, bl & 31);
                    if ((lt & 31) == 0) { hd[qtr] = hd0; td[qtr] = has_tail ? tdv : -1; }
                } else {
                    // several boundaries inside 32 rows (graphs of fewer than 34 vertices): general scan
                    Stat cur;
                    stat_init(cur);
                    int32_t cd = hd0;
                    bool seen = false;
#pragma unroll 1
                    for (int g = 0; g < 4; ++g) {
                        const int rb = r0 + 8 * g;
                        if (rb >= r1) break;
                        float v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] = *reinterpret_cast<const float*>(base + (rb + k) * 128 + ((cc ^ k) << 4));
#pragma unroll 1
                        for (int k = 0; k < 8; ++k) {
                            const int r = rb + k;
                            if (r < r1) {
                                if ((starts >> (8 * g + k)) & 1u) {
                                    if (!seen) {
                                        S.red.f[0][0][qtr][col] = cur.sum; S.red.f[0][1][qtr][col] = cur.sq;
                                        S.red.f[0][2][qtr][col] = cur.mn;  S.red.f[0][3][qtr][col] = cur.mx;
                                        S.red.a[0][0][qtr][col] = cur.amn; S.red.a[0][1][qtr][col] = cur.amx;
                                        seen = true;
                                    } else {
                                        agg_write_ool(a.out, cd, col, cur);
                                    }
                                    stat_init(cur);
                                    cd = __shfl_sync(0xffffffffu, d, 8 * g + k);
                                }
                                const float vk = k == 0 ? v[0] : k == 1 ? v[1] : k == 2 ? v[2] : k == 3 ? v[3] : k == 4 ? v[4] : k == 5 ? v[5] : k == 6 ? v[6] : v[7];
                                if (WANT_ARG) stat_add(cur, vk, t0 + r);
                                else stat_add_noarg(cur, vk);
                            }
                        }
                    }
                    const int w = seen ? 1 : 0;
                    S.red.f[w][0][qtr][col] = cur.sum; S.red.f[w][1][qtr][col] = cur.sq;
                    S.red.f[w][2][qtr][col] = cur.mn;  S.red.f[w][3][qtr][col] = cur.mx;
                    S.red.a[w][0][qtr][col] = cur.amn; S.red.a[w][1][qtr][col] = cur.amx;
                    if ((lt & 31) == 0) { hd[qtr] = hd0; td[qtr] = seen ? cd : -1; }
                }
            }
            if (has_out && nrows < TILE) {         // partial last tile of the range: plain stores of this slab, nothing past e1
                for (int idx = lt; idx < nrows * 8; idx += 128) {
                    const int r = idx >> 3, c4l = idx & 7;
                    __stcs(reinterpret_cast<float4*>(a.m_out + (int64_t)(t0 + r) * 64 + 32 * grp) + c4l,
                           *reinterpret_cast<const float4*>(l_ptr(S.M, r, 8 * grp + c4l)));
                }
            }
            FW_TR(lt == 0 && grp == 0, 6, i, 1);
            warp_arrive(&S.bar_Mfree[grp], lane);  // this warp is done reading the slab
            named_sync(3 + grp, 128);              // all head / tail partials of the group are in the scratch
            if (lt < 32) {
                // carry chain over the four quarters (same logic as tile_reduce_chain, segreduce.cuh)
                for (int qq = 0; qq < 4; ++qq) {
                    if (qq * 32 >= nrows) break;
                    Stat h;
                    h.sum = S.red.f[0][0][qq][col]; h.sq = S.red.f[0][1][qq][col];
                    h.mn = S.red.f[0][2][qq][col];  h.mx = S.red.f[0][3][qq][col];
                    h.amn = S.red.a[0][0][qq][col]; h.amx = S.red.a[0][1][qq][col];
                    const int32_t hdq = hd[qq];
                    if (carry.dst == hdq) {
                        stat_merge(carry.s, h);
                    } else {
                        if (carry.dst >= 0) agg_write_ool(a.out, carry.dst, col, carry.s);
                        carry.s = h; carry.dst = hdq;
                    }
                    const int32_t tdq = td[qq];
                    if (tdq >= 0) {
                        agg_write_ool(a.out, carry.dst, col, carry.s);
                        carry.s.sum = S.red.f[1][0][qq][col]; carry.s.sq = S.red.f[1][1][qq][col];
                        carry.s.mn = S.red.f[1][2][qq][col];  carry.s.mx = S.red.f[1][3][qq][col];
                        carry.s.amn = S.red.a[1][0][qq][col]; carry.s.amx = S.red.a[1][1][qq][col];
                        carry.dst = tdq;
                    }
                }
            }
        }
        if (lt < 32) carry_flush_t(carry, a.out, col);
    }
    tc_fence_before();
    __syncthreads();
    if (a.trace && tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        a.trace[7 * 64 * 8 + blockIdx.x] = ((long long)smid << 48) | (clock64() - t_start);
    }
    if (warp == 22) tmem_dealloc<512>(tmem);
}

template <bool FIRST, bool X3, bool WANT_ARG>
static int launch_fwd_ws_variant(const CUtensorMap& tin, const CUtensorMap& tout, const FwdWsArgs& a, int grid, cudaStream_t st) {
    const int smem = (int)sizeof(FwdWsSmem) + 1024;
    static bool done = false;
    if (!done) {
        GCRL_CUDA(cudaFuncSetAttribute(k_edge_fwd_ws<FIRST, X3, WANT_ARG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        done = true;
    }
    static int n_launch = 0;
    const char* tr = getenv("GCRL_WS_TRACE_FWD");
    FwdWsArgs b = a;
    long long* d_tr = nullptr;
    const size_t tr_n = 7 * 64 * 8 + 160;
    if (tr && !FIRST && ++n_launch == 3) {
        GCRL_CUDA(cudaMalloc(&d_tr, tr_n * sizeof(long long)));
        GCRL_CUDA(cudaMemsetAsync(d_tr, 0, tr_n * sizeof(long long), st));
        b.trace = d_tr;
    }
    k_edge_fwd_ws<FIRST, X3, WANT_ARG><<<grid, FW_THREADS, smem, st>>>(tin, tout, b);
    if (d_tr) {
        long long* h = (long long*)malloc(tr_n * sizeof(long long));
        GCRL_CUDA(cudaStreamSynchronize(st));
        GCRL_CUDA(cudaMemcpy(h, d_tr, tr_n * sizeof(long long), cudaMemcpyDeviceToHost));
        FILE* f = fopen(tr, "w");
        if (f) {
            for (size_t i = 0; i < 7 * 64 * 8; ++i)
                if (h[i]) fprintf(f, "%d %d %d %lld\n", (int)(i / 512), (int)((i / 8) % 64), (int)(i % 8), h[i]);
            for (size_t i = 7 * 64 * 8; i < tr_n; ++i)
                if (h[i]) fprintf(f, "cta %d sm %d cycles %lld\n", (int)(i - 7 * 64 * 8), (int)(h[i] >> 48), h[i] & 0xffffffffffffll);
            fclose(f);
        }
        free(h);
        cudaFree(d_tr);
    }
    return GCRL_OK;
}

// GCRL_FWD_IMPL=ws selects this kernel.  It is parity-green but, in bf16x3, not yet faster than the single-group
// kernel with two CTAs per SM (B200, 256 x 200-vertex graphs: 2.8 ms vs 2.4 ms per launch): the ALU-heavy reducer
// warps and the latency-bound gather group share issue slots, and the gather group paces the MMA chain.
bool use_ws_forward() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GCRL_FWD_IMPL");
        v = (e && e[0] == 'w' && e[1] == 's') ? 1 : 0;
    }
    return v == 1;
}

// Warp-specialised tensor-core edge forward; same contract as launch_edge_fwd_tc (gn_forward_tc.cu).
int launch_edge_fwd_ws(const BlockW& b, bool first, bool x3, const float* Pi, const float* Pj, const float* e_prev,
                       const int32_t* seg_off, const int32_t* src, const int32_t* dst, int64_t N, int64_t E,
                       float* m_out, float* agg, float* var, int32_t* amin, int32_t* amax, cudaStream_t st) {
    if (E == 0) return GCRL_OK;
    FwdWsArgs a;
    a.Pi = Pi; a.Pj = Pj; a.e_prev = e_prev; a.De = b.De;
    a.W1 = b.W1; a.ldW1 = 2 * b.Dv + b.De; a.coff = 2 * b.Dv;
    a.W2 = b.W2; a.b2 = b.b2;
    a.seg_off = seg_off; a.src = src; a.dst = dst; a.N = (int32_t)N; a.E = (int32_t)E;
    a.m_out = m_out;
    a.out.agg = agg; a.out.var = var; a.out.amin = amin; a.out.amax = amax; a.out.seg_off = seg_off;
    a.trace = nullptr;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("GCRL_FW_DBG"); dbg = e ? atoi(e) : 0; } a.dbg = dbg; }
    CUtensorMap tin, tout;
    int rc = make_tmap_rows64(&tin, first ? Pi : e_prev, first ? N : E, 32);
    if (rc) return rc;
    rc = make_tmap_rows64(&tout, m_out ? m_out : Pi, m_out ? E : N, 128);
    if (rc) return rc;
    int64_t grid = sm_count();
    const int64_t tiles = (E + TILE - 1) / TILE;
    if (grid > tiles) grid = tiles;
    const int tok = prof_begin(first ? GCRL_PROF_EDGE_FWD_FIRST : (m_out ? GCRL_PROF_EDGE_FWD : GCRL_PROF_EDGE_FWD_LAST), st);
    const bool want_arg = amin != nullptr;
#define FW_LAUNCH(F, X, W) rc = launch_fwd_ws_variant<F, X, W>(tin, tout, a, (int)grid, st)
    if (first) { if (x3) { if (want_arg) FW_LAUNCH(true, true, true); else FW_LAUNCH(true, true, false); }
                 else    { if (want_arg) FW_LAUNCH(true, false, true); else FW_LAUNCH(true, false, false); } }
    else       { if (x3) { if (want_arg) FW_LAUNCH(false, true, true); else FW_LAUNCH(false, true, false); }
                 else    { if (want_arg) FW_LAUNCH(false, false, true); else FW_LAUNCH(false, false, false); } }
#undef FW_LAUNCH
    prof_end(tok, st);
    if (rc) return rc;
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

}  // namespace gcrl
