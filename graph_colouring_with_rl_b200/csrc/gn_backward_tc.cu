// gn_backward_tc.cu -- the fused edge kernel of the GN backward pass on the tensor cores.
//
// Same math as k_edge_bwd (gn_backward.cu) per 128-edge tile, with every product on tcgen05
// (kind::f16, bf16 hi/lo split operands, fp32 accumulators in TMEM):
//
//   G1  acc1  = e_{l-1} . C^T                     recompute of the hidden pre-activation
//   G2  acc2  = hid . W2^T                        block 5 only (its e_5 was never stored)
//   G3  accW2 += dm^T . hid                       weight gradient, contracts over the 128 edge rows
//   G4  acc4  = dm . W2                           d hid
//   G5  accC  += dz1^T . e_{l-1}                  weight gradient of the edge columns of W1
//   G6  acc6  = dz1 . C                           d e_{l-1}
//   G7  accPi = S^T . dz1                         per-destination sums (S = row -> segment one-hot)
//
// The bf16 operand tile format is simultaneously K-major and MN-major (tc.cuh), so dm, dz1, hid
// and e_{l-1} are written to shared memory ONCE and feed both the feature-contracting products
// (G1,G2,G4,G6) and the row-contracting ones (G3,G5,G7); W2 and C are likewise stored once and
// read K-major (forward products) or MN-major (their transposes).  The two weight-gradient
// accumulators stay in TMEM for the whole kernel and are flushed with one atomic pass per CTA.
//
// Data movement: TMA loads e_{l-1}, e_l and d e_l tiles (fp32, SWIZZLE_128B) and stores d e_{l-1};
// dPj[src] += dz1 goes out as 128-bit vector reds from registers; dPi[dst] += (segment sums of
// dz1) comes from G7 -- a handful of rows per tile instead of 128.
// PNA backward uses per-vertex coefficient arrays (k_pna_bwd_coef):
//   dm = d e_l + ca[v] + cb[v] * m + [slot == argmin] g_min + [slot == argmax] g_max.
#define GCRL_MBAR_TIMEOUT   // counted waits (with a trap on protocol bugs): measurably faster here than the tight loop
#include "gn_common.cuh"
#include "tc.cuh"

namespace gcrl {

using namespace tc;

// ca = g_mean/deg - [var>0] g_std * mean / (deg * std),  cb = [var>0] g_std / (deg * std)
// With de_scatter != null the min / max gradient terms are added to d e_l here (two scattered reds per
// (vertex, column): 1/(n-1) of the elements) so that the edge kernel's dm pass is purely dense:
//   d e_l[argmin[v,c], c] += g_min[v,c],  d e_l[argmax[v,c], c] += g_max[v,c].
__global__ void k_pna_bwd_coef(const float* __restrict__ dagg, const float* __restrict__ agg,
                               const float* __restrict__ var, const int32_t* __restrict__ seg_off, int64_t N,
                               float* __restrict__ ca, float* __restrict__ cb, const int32_t* __restrict__ amin,
                               const int32_t* __restrict__ amax, float* __restrict__ de_scatter,
                               float* __restrict__ gb2_vertex) {
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

struct BwdTcSmem {
    __nv_bfloat16 Wc[2][64 * 64];    // C hi/lo (rows = o, cols = k): K-major for G1, MN-major for G6
    __nv_bfloat16 W2[2][64 * 64];    // W2 hi/lo: K-major for G2, MN-major for G4
    float L[128 * 64];               // fp32 slab tile: e_{l-1} landing, then d e_l landing
    float Lx[128 * 64];              // fp32 slab tile: e_l landing, then d e_{l-1} staging
    __nv_bfloat16 B1[2][128 * 64];   // e_{l-1} operand
    __nv_bfloat16 B2[2][128 * 64];   // hid operand
    __nv_bfloat16 B3[2][128 * 64];   // Pi+Pj (fp32 view), then dm, then dz1
    __nv_bfloat16 IND[128 * 64];     // row -> segment one-hot (bf16 1.0)
    float b2[64];
    float bsum[64];                  // final db2 reduction
    int32_t dst[TILE];
    int32_t src[TILE];
    int32_t seg_dst[64];
    uint32_t start_bits[4];
    uint64_t bar_e, bar_m, bar_de, bar_mma;
    uint32_t tmem_base;
};

struct BwdTcArgs {
    const float* Pi; const float* Pj;
    const float* e_prev; int De;
    const float* W1; int ldW1, coff;
    const float* W2; const float* b2;
    const int32_t* seg_off; const int32_t* src; const int32_t* dst;
    int32_t N, E;
    const float* ca; const float* cb;       // [N,64] PNA coefficients
    const float* dagg;                      // [N,256] (min / max columns read here)
    const int32_t* amin; const int32_t* amax;
    float* de_prev;                         // manual store path of the last partial tile
    float* dPi; float* dPj;
    float* gW1; float* gW2; float* gb2;
};

template <bool FIRST, bool LAST, bool X3>
__global__ void __launch_bounds__(256, 1) k_edge_bwd_tc(const __grid_constant__ CUtensorMap tm_e,
                                                       const __grid_constant__ CUtensorMap tm_m,
                                                       const __grid_constant__ CUtensorMap tm_de,
                                                       const __grid_constant__ CUtensorMap tm_out, const BwdTcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    BwdTcSmem& S = *reinterpret_cast<BwdTcSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sp = warp & 3, half = warp >> 2;
    const int row = 32 * sp + lane;
    float* P3 = reinterpret_cast<float*>(&S.B3[0][0]);        // fp32 slab view of B3 (Pi + Pj)

    if (tid == 0) {
        mbar_init(&S.bar_e, 1);
        mbar_init(&S.bar_m, 1);
        mbar_init(&S.bar_de, 1);
        mbar_init(&S.bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(&S.tmem_base);
    load_weight_tile(a.W2, 64, 0, S.W2[0], S.W2[1], tid, 256);
    if (!FIRST) load_weight_tile(a.W1, a.ldW1, a.coff, S.Wc[0], S.Wc[1], tid, 256);
    float (*Cf)[64] = reinterpret_cast<float (*)[64]>(&S.Wc[0][0]);   // FIRST: C^T fp32 [d][o]
    if (FIRST) {
        for (int t = tid; t < a.De * 64; t += 256) Cf[t >> 6][t & 63] = a.W1[(t & 63) * a.ldW1 + a.coff + (t >> 6)];
        for (int t = tid; t < 2 * 128 * 8; t += 256)   // zero-padded e tile
            reinterpret_cast<uint4*>(&S.B1[0][0])[t] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int t = tid; t < 128 * 8; t += 256) reinterpret_cast<uint4*>(S.IND)[t] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 64) { S.b2[tid] = a.b2[tid]; S.bsum[tid] = 0.f; }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    const uint32_t t_acc1 = tmem, t_acc2 = tmem + 64, t_acc4 = tmem + 128, t_acc6 = tmem + 192;
    const uint32_t t_w2 = tmem + 256, t_c = tmem + 320, t_pi = tmem + 384;
    const uint32_t lane_addr = (uint32_t)(32 * sp) << 16;
    const uint32_t id_feat = make_idesc_bf16(128, 64, false, false);   // K-major A, K-major B
    const uint32_t id_featT = make_idesc_bf16(128, 64, false, true);   // K-major A, MN-major B (weight transposed)
    const uint32_t id_rows = make_idesc_bf16(64, 64, true, true);      // MN-major A and B (contract over rows)
    uint32_t ph_e = 0, ph_m = 0, ph_de = 0, ph_mma = 0;

    float bsum[32];                 // db2 partial: this thread's rows x its 32 columns
#pragma unroll
    for (int j = 0; j < 32; ++j) bsum[j] = 0.f;
    int prev_seg = -1;              // tid < 128: column of this row's 1 in IND

    const int64_t n_tiles = ((int64_t)a.E + TILE - 1) / TILE;
    int64_t tile = blockIdx.x;
    if (tile < n_tiles && tid == 0) {
        const int32_t t0 = (int32_t)(tile * TILE);
        if (!FIRST) {
            mbar_arrive_expect_tx(&S.bar_e, 128 * 64 * 4);
            tma_load_2d(S.L, &tm_e, 0, t0, &S.bar_e);
            tma_load_2d(S.L + 128 * 32, &tm_e, 32, t0, &S.bar_e);
        }
        if (!LAST) {
            mbar_arrive_expect_tx(&S.bar_m, 128 * 64 * 4);
            tma_load_2d(S.Lx, &tm_m, 0, t0, &S.bar_m);
            tma_load_2d(S.Lx + 128 * 32, &tm_m, 32, t0, &S.bar_m);
        }
    }
    int32_t nd = -1, ns = 0;
    if (tile < n_tiles && tid < TILE && tile * TILE + tid < a.E) {
        nd = __ldg(a.dst + tile * TILE + tid);
        ns = __ldg(a.src + tile * TILE + tid);
    }
    bool first_tile = true;
    for (; tile < n_tiles; tile += gridDim.x) {
        const int32_t t0 = (int32_t)(tile * TILE);
        const int nrows = min(TILE, a.E - t0);
        const int64_t next = tile + gridDim.x;
        const bool has_next = next < n_tiles;
        if (tid < TILE) { S.dst[tid] = nd; S.src[tid] = ns; }
        __syncthreads();                                                                   // #1
        bool my_start = false;
        uint32_t my_bits = 0;
        if (tid < TILE) {
            my_start = tid < nrows && (tid == 0 || S.dst[tid] != S.dst[tid - 1]);
            my_bits = __ballot_sync(0xffffffffu, my_start);
            if (lane == 0) S.start_bits[warp] = my_bits;
            nd = -1; ns = 0;
            if (has_next && next * TILE + tid < a.E) {
                nd = __ldg(a.dst + next * TILE + tid);
                ns = __ldg(a.src + next * TILE + tid);
            }
        }
        // Pi[dst] + Pj[src] row gathers (coalesced: half a warp per row); consumed after the operand split
        float4 pg[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int t = it * 256 + tid;
            const int r = t >> 4, c4 = t & 15;
            pg[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrows) {
                const float4 pi = __ldg(reinterpret_cast<const float4*>(a.Pi + (int64_t)S.dst[r] * 64) + c4);
                const float4 pj = __ldg(reinterpret_cast<const float4*>(a.Pj + (int64_t)S.src[r] * 64) + c4);
                pg[it] = make_float4(pi.x + pj.x, pi.y + pj.y, pi.z + pj.z, pi.w + pj.w);
            }
        }
        if (!FIRST) {
            mbar_wait(&S.bar_e, ph_e);
            ph_e ^= 1;
            convert_tile128<X3>(S.L, S.B1[0], S.B1[1], tid);
        } else if (tid < TILE) {
            // zero-padded raw edge features: chunk 0 of the row holds e[0..De)
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = (tid < nrows && j < a.De) ? a.e_prev[(int64_t)(t0 + tid) * a.De + j] : 0.f;
            store_split8<X3>(x, S.B1[0], S.B1[1], tid, 0);
        }
        // previous tile's G5/G6/G7 (readers of B3) completed before its epilogue 4: B3 is free
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int t = it * 256 + tid;
            *reinterpret_cast<float4*>(l_ptr(P3, t >> 4, t & 15)) = pg[it];
        }
        fence_proxy_async();
        __syncthreads();                                                                   // #2
        if (warp == 0 && elect_one()) {   // one elected lane of a converged warp: no per-thread election loop around each UMMA
            tc_fence_after();
            if (!FIRST) {
                issue_gemm_feat<X3>(t_acc1, smem_u32(S.B1[0]), smem_u32(S.B1[1]), smem_u32(S.Wc[0]), smem_u32(S.Wc[1]), id_feat, false);
                umma_commit(&S.bar_mma);
            }
            if (!LAST) {                       // L is free after the split: bring in d e_l
                mbar_arrive_expect_tx(&S.bar_de, 128 * 64 * 4);
                tma_load_2d(S.L, &tm_de, 0, t0, &S.bar_de);
                tma_load_2d(S.L + 128 * 32, &tm_de, 32, t0, &S.bar_de);
            }
        }
        // ---- segment bookkeeping: one-hot row -> segment tile for G7 --------------------------------------
        int n_segs = 0, seg_of_row = -1;
        {
            int before = 0;
#pragma unroll
            for (int w2 = 0; w2 < 4; ++w2) {
                const int pc = __popc(S.start_bits[w2]);
                if (w2 < sp) before += pc;
                n_segs += pc;
            }
            if (row < nrows) seg_of_row = before + __popc(S.start_bits[sp] & ((2u << lane) - 1u)) - 1;
        }
        if (tid < TILE) {     // warps 0-3: row == tid
            if (my_start && seg_of_row < 64) S.seg_dst[seg_of_row] = S.dst[tid];
            const int new_col = (seg_of_row >= 0 && seg_of_row < 64) ? seg_of_row : -1;
            if (new_col != prev_seg) {
                unsigned char* line = reinterpret_cast<unsigned char*>(S.IND);
                if (prev_seg >= 0) *reinterpret_cast<uint4*>(line + bf_off(tid, prev_seg >> 3)) = make_uint4(0u, 0u, 0u, 0u);
                if (new_col >= 0) {
                    const int wi = (new_col & 7) >> 1;
                    const uint32_t one = (new_col & 1) ? 0x3F800000u : 0x00003F80u;   // bf16 1.0 in the element's half
                    *reinterpret_cast<uint4*>(line + bf_off(tid, new_col >> 3)) =
                        make_uint4(wi == 0 ? one : 0u, wi == 1 ? one : 0u, wi == 2 ? one : 0u, wi == 3 ? one : 0u);
                }
                prev_seg = new_col;
            }
        }

        // ---- epilogue 1: hid = relu(acc1 + Pi + Pj) -> B2 ; relu mask kept in a register -------------------------
        float acc[32];
        if (!FIRST) {
            mbar_wait(&S.bar_mma, ph_mma);
            ph_mma ^= 1;
            tc_fence_after();
            tmem_ld32(t_acc1 + lane_addr + 32 * half, acc);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0.f;
            if (row < nrows) {
                for (int d = 0; d < a.De; ++d) {
                    const float ev = a.e_prev[(int64_t)(t0 + row) * a.De + d];
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = fmaf(ev, Cf[d][32 * half + j], acc[j]);
                }
            }
        }
        uint32_t hmask = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float x[8];
            const float4 p0 = *reinterpret_cast<const float4*>(l_ptr(P3, row, 8 * half + 2 * c));
            const float4 p1 = *reinterpret_cast<const float4*>(l_ptr(P3, row, 8 * half + 2 * c + 1));
            x[0] = fmaxf(p0.x + acc[8 * c + 0], 0.f); x[1] = fmaxf(p0.y + acc[8 * c + 1], 0.f);
            x[2] = fmaxf(p0.z + acc[8 * c + 2], 0.f); x[3] = fmaxf(p0.w + acc[8 * c + 3], 0.f);
            x[4] = fmaxf(p1.x + acc[8 * c + 4], 0.f); x[5] = fmaxf(p1.y + acc[8 * c + 5], 0.f);
            x[6] = fmaxf(p1.z + acc[8 * c + 6], 0.f); x[7] = fmaxf(p1.w + acc[8 * c + 7], 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) hmask |= (x[j] > 0.f ? 1u : 0u) << (8 * c + j);
            store_split8<X3>(x, S.B2[0], S.B2[1], row, 4 * half + c);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();                                                                   // #3
        if (LAST) {
            if (warp == 0 && elect_one()) {   // one elected lane of a converged warp: no per-thread election loop around each UMMA
                tc_fence_after();
                issue_gemm_feat<X3>(t_acc2, smem_u32(S.B2[0]), smem_u32(S.B2[1]), smem_u32(S.W2[0]), smem_u32(S.W2[1]), id_feat, false);
                umma_commit(&S.bar_mma);
            }
            mbar_wait(&S.bar_mma, ph_mma);
            ph_mma ^= 1;
            tc_fence_after();
            tmem_ld32(t_acc2 + lane_addr + 32 * half, acc);
        } else {
            mbar_wait(&S.bar_m, ph_m);
            ph_m ^= 1;
            mbar_wait(&S.bar_de, ph_de);
            ph_de ^= 1;
        }
        // ---- epilogue 2: dm = d e_l + ca + cb * m + arg terms -> B3 -------------------------------------------------
        {
            const bool valid = row < nrows;
            const int32_t v = valid ? S.dst[row] : 0;
            const int32_t slot = t0 + row;
            const float4* ca4 = reinterpret_cast<const float4*>(a.ca + (int64_t)v * 64 + 32 * half);
            const float4* cb4 = reinterpret_cast<const float4*>(a.cb + (int64_t)v * 64 + 32 * half);
            const float4* gmn4 = reinterpret_cast<const float4*>(a.dagg + (int64_t)v * 256 + 64 + 32 * half);
            const float4* gmx4 = reinterpret_cast<const float4*>(a.dagg + (int64_t)v * 256 + 128 + 32 * half);
            const int4* imn4 = reinterpret_cast<const int4*>(a.amin + (int64_t)v * 64 + 32 * half);
            const int4* imx4 = reinterpret_cast<const int4*>(a.amax + (int64_t)v * 64 + 32 * half);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float x[8];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int q4 = 2 * c + hh;          // float4 index inside this thread's 32 columns
                    float4 m4, d4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (LAST) {
                        const float4 bb = *reinterpret_cast<const float4*>(&S.b2[32 * half + 4 * q4]);
                        m4 = make_float4(acc[4 * q4] + bb.x, acc[4 * q4 + 1] + bb.y, acc[4 * q4 + 2] + bb.z, acc[4 * q4 + 3] + bb.w);
                    } else {
                        m4 = *reinterpret_cast<const float4*>(l_ptr(S.Lx, row, 8 * half + q4));
                        d4 = *reinterpret_cast<const float4*>(l_ptr(S.L, row, 8 * half + q4));
                    }
                    float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid) {
                        const float4 A4 = __ldg(ca4 + q4), B4 = __ldg(cb4 + q4);
                        const float4 gmn = __ldg(gmn4 + q4), gmx = __ldg(gmx4 + q4);
                        const int4 imn = __ldg(imn4 + q4), imx = __ldg(imx4 + q4);
                        r4.x = fmaf(B4.x, m4.x, A4.x) + d4.x; r4.y = fmaf(B4.y, m4.y, A4.y) + d4.y;
                        r4.z = fmaf(B4.z, m4.z, A4.z) + d4.z; r4.w = fmaf(B4.w, m4.w, A4.w) + d4.w;
                        if (imn.x == slot) r4.x += gmn.x;
                        if (imn.y == slot) r4.y += gmn.y;
                        if (imn.z == slot) r4.z += gmn.z;
                        if (imn.w == slot) r4.w += gmn.w;
                        if (imx.x == slot) r4.x += gmx.x;
                        if (imx.y == slot) r4.y += gmx.y;
                        if (imx.z == slot) r4.z += gmx.z;
                        if (imx.w == slot) r4.w += gmx.w;
                    }
                    x[4 * hh] = r4.x; x[4 * hh + 1] = r4.y; x[4 * hh + 2] = r4.z; x[4 * hh + 3] = r4.w;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) bsum[8 * c + j] += x[j];
                store_split8<X3>(x, S.B3[0], S.B3[1], row, 4 * half + c);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();                                                                   // #4
        if (warp == 0 && elect_one()) {   // one elected lane of a converged warp: no per-thread election loop around each UMMA
            tc_fence_after();
            // G3: dW2 += dm^T hid ; G4: d hid = dm W2
            issue_gemm_rows<X3>(t_w2, smem_u32(S.B3[0]), smem_u32(S.B3[1]), smem_u32(S.B2[0]), smem_u32(S.B2[1]), id_rows, 128, !first_tile);
            issue_gemm_featT<X3>(t_acc4, smem_u32(S.B3[0]), smem_u32(S.B3[1]), smem_u32(S.W2[0]), smem_u32(S.W2[1]), id_featT, false);
            umma_commit(&S.bar_mma);
            if (!FIRST && has_next) {              // L consumed by epilogue 2: prefetch the next tile's e_{l-1}
                mbar_arrive_expect_tx(&S.bar_e, 128 * 64 * 4);
                tma_load_2d(S.L, &tm_e, 0, (int32_t)(next * TILE), &S.bar_e);
                tma_load_2d(S.L + 128 * 32, &tm_e, 32, (int32_t)(next * TILE), &S.bar_e);
            }
        }
        mbar_wait(&S.bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        // ---- epilogue 3: dz1 = d hid * relu' -> B3 ; dPj[src] += dz1 ---------------------------------------------------
        tmem_ld32(t_acc4 + lane_addr + 32 * half, acc);
        {
            const bool valid = row < nrows;
            float4* pj4 = reinterpret_cast<float4*>(a.dPj + (int64_t)(valid ? S.src[row] : 0) * 64 + 32 * half);
            float4* pi4 = reinterpret_cast<float4*>(a.dPi + (int64_t)(valid ? S.dst[row] : 0) * 64 + 32 * half);
            const bool overflow_seg = valid && seg_of_row >= 64;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = ((hmask >> (8 * c + j)) & 1u) ? acc[8 * c + j] : 0.f;
                if (valid) {
                    atomicAdd(pj4 + 2 * c, make_float4(x[0], x[1], x[2], x[3]));
                    atomicAdd(pj4 + 2 * c + 1, make_float4(x[4], x[5], x[6], x[7]));
                    if (overflow_seg) {      // more than 64 segments in this tile (graphs of <= 2 vertices)
                        atomicAdd(pi4 + 2 * c, make_float4(x[0], x[1], x[2], x[3]));
                        atomicAdd(pi4 + 2 * c + 1, make_float4(x[4], x[5], x[6], x[7]));
                    }
                }
                store_split8<X3>(x, S.B3[0], S.B3[1], row, 4 * half + c);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();                                                                   // #5
        if (warp == 0 && elect_one()) {   // one elected lane of a converged warp: no per-thread election loop around each UMMA
            tc_fence_after();
            // G5: dC += dz1^T e_{l-1} ; G6: d e_{l-1} = dz1 C ; G7: per-segment sums of dz1
            issue_gemm_rows<X3>(t_c, smem_u32(S.B3[0]), smem_u32(S.B3[1]), smem_u32(S.B1[0]), smem_u32(S.B1[1]), id_rows, 128, !first_tile);
            if (!FIRST)
                issue_gemm_featT<X3>(t_acc6, smem_u32(S.B3[0]), smem_u32(S.B3[1]), smem_u32(S.Wc[0]), smem_u32(S.Wc[1]), id_featT, false);
            issue_gemm_rows<false>(t_pi, smem_u32(S.IND), 0, smem_u32(S.B3[0]), 0, id_rows, 128, false);
            if (X3) issue_gemm_rows<false>(t_pi, smem_u32(S.IND), 0, smem_u32(S.B3[1]), 0, id_rows, 128, true);
            umma_commit(&S.bar_mma);
        }
        mbar_wait(&S.bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        // ---- epilogue 4: d e_{l-1} -> Lx -> TMA store ; dPi[dst] += segment sums ---------------------------------------
        if (!FIRST) {
            tmem_ld32(t_acc6 + lane_addr + 32 * half, acc);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(l_ptr(S.Lx, row, 8 * half + c)) = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
            fence_proxy_async();
        }
        if (warp < 4) {
            // accPi row s (segment) lives in lane (s % 16) + 32 * (s / 16): warp w owns segments 16w .. 16w+15
            const int s = 16 * warp + lane;
            const bool has = lane < 16 && s < n_segs && s < 64;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                tmem_ld32(t_pi + lane_addr + 32 * hh, acc);
                if (has) {
                    float4* o4 = reinterpret_cast<float4*>(a.dPi + (int64_t)S.seg_dst[s] * 64 + 32 * hh);
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        atomicAdd(o4 + c, make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]));
                }
            }
        }
        tc_fence_before();
        __syncthreads();                                                                   // #6
        if (!FIRST) {
            if (nrows == TILE) {
                if (tid == 0) {
                    tma_store_2d(&tm_out, 0, t0, S.Lx);
                    tma_store_2d(&tm_out, 32, t0, S.Lx + 128 * 32);
                    bulk_commit();
                    bulk_wait_read0();
                }
            } else {
                float4* g = reinterpret_cast<float4*>(a.de_prev + (int64_t)t0 * 64);
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int idx = it * 256 + tid;
                    const int r = idx >> 4, c4 = idx & 15;
                    if (r < nrows) __stcs(g + idx, *reinterpret_cast<const float4*>(l_ptr(S.Lx, r, c4)));
                }
                __syncthreads();
            }
        }
        if (!LAST && has_next && tid == 0) {        // Lx is free again: next tile's e_l
            mbar_arrive_expect_tx(&S.bar_m, 128 * 64 * 4);
            tma_load_2d(S.Lx, &tm_m, 0, (int32_t)(next * TILE), &S.bar_m);
            tma_load_2d(S.Lx + 128 * 32, &tm_m, 32, (int32_t)(next * TILE), &S.bar_m);
        }
        first_tile = false;
    }

    // ---- flush: weight gradients from TMEM, bias gradient from registers -----------------------------------------------
    if (!first_tile) {
        tc_fence_after();
        if (warp < 4) {
            const int o = 16 * warp + lane;          // accumulator row (lanes 0..15 of each sub-partition)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float v[32];
                tmem_ld32(t_w2 + lane_addr + 32 * hh, v);
                if (lane < 16)
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(&a.gW2[o * 64 + 32 * hh + j], v[j]);
                tmem_ld32(t_c + lane_addr + 32 * hh, v);
                if (lane < 16) {
                    if (!FIRST) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(&a.gW1[o * a.ldW1 + a.coff + 32 * hh + j], v[j]);
                    } else if (hh == 0) {
                        // FIRST: accC[o][d] = sum_r dz1[r][o] e[r][d] for d < De (zero-padded tile)
                        for (int j = 0; j < a.De; ++j) atomicAdd(&a.gW1[o * a.ldW1 + a.coff + j], v[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(&S.bsum[32 * half + j], bsum[j]);
        __syncthreads();
        if (tid < 64) atomicAdd(&a.gb2[tid], S.bsum[tid]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

template <bool FIRST, bool LAST, bool X3>
static int launch_bwd_variant(const CUtensorMap& te, const CUtensorMap& tm, const CUtensorMap& tde, const CUtensorMap& tout,
                              const BwdTcArgs& a, int grid, cudaStream_t st) {
    const int smem = (int)sizeof(BwdTcSmem) + 1024;
    static bool done = false;
    if (!done) {
        GCRL_CUDA(cudaFuncSetAttribute(k_edge_bwd_tc<FIRST, LAST, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        done = true;
    }
    k_edge_bwd_tc<FIRST, LAST, X3><<<grid, 256, smem, st>>>(te, tm, tde, tout, a);
    return GCRL_OK;
}

int launch_pna_bwd_coef(const float* dagg, const float* agg, const float* var, const int32_t* seg_off, int64_t N,
                        float* ca, float* cb, const int32_t* amin, const int32_t* amax, float* de_scatter,
                        float* gb2_vertex, cudaStream_t st) {
    if (N == 0) return GCRL_OK;
    int64_t blocks = (N * 64 + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_pna_bwd_coef<<<(int)blocks, 256, 0, st>>>(dagg, agg, var, seg_off, N, ca, cb, amin, amax, de_scatter, gb2_vertex);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

// Tensor-core edge backward of one block.  m == nullptr <=> last block (e_5 recomputed, no d e_l).
int launch_edge_bwd_tc(bool first, bool x3, const float* Pi, const float* Pj, const float* e_prev, int De,
                       const float* W1, int ldW1, int coff, const float* W2, const float* b2, const int32_t* seg_off,
                       const int32_t* src, const int32_t* dst, int64_t N, int64_t E, const float* m, const float* de,
                       const float* ca, const float* cb, const float* dagg, const int32_t* amin, const int32_t* amax,
                       float* de_prev, float* dPi, float* dPj, float* gW1, float* gW2, float* gb2, cudaStream_t st) {
    if (E == 0) return GCRL_OK;
    const bool last = (m == nullptr);
    BwdTcArgs a;
    a.Pi = Pi; a.Pj = Pj; a.e_prev = e_prev; a.De = De;
    a.W1 = W1; a.ldW1 = ldW1; a.coff = coff; a.W2 = W2; a.b2 = b2;
    a.seg_off = seg_off; a.src = src; a.dst = dst; a.N = (int32_t)N; a.E = (int32_t)E;
    a.ca = ca; a.cb = cb; a.dagg = dagg; a.amin = amin; a.amax = amax;
    a.de_prev = de_prev; a.dPi = dPi; a.dPj = dPj; a.gW1 = gW1; a.gW2 = gW2; a.gb2 = gb2;
    CUtensorMap te, tm, tde, tout;
    int rc;
    // unused maps point at a valid [N,64] array so that encoding never fails
    if ((rc = make_tmap_rows64(&te, first ? Pi : e_prev, first ? N : E, 128))) return rc;
    if ((rc = make_tmap_rows64(&tm, last ? Pi : m, last ? N : E, 128))) return rc;
    if ((rc = make_tmap_rows64(&tde, last ? Pi : de, last ? N : E, 128))) return rc;
    if ((rc = make_tmap_rows64(&tout, first ? Pi : de_prev, first ? N : E, 128))) return rc;
    int64_t tiles = (E + TILE - 1) / TILE;
    int64_t grid = sm_count();
    if (grid > tiles) grid = tiles;
    const int tok = prof_begin(first ? GCRL_PROF_EDGE_BWD_FIRST : (last ? GCRL_PROF_EDGE_BWD_LAST : GCRL_PROF_EDGE_BWD), st);
    if (first) rc = x3 ? launch_bwd_variant<true, false, true>(te, tm, tde, tout, a, (int)grid, st)
                       : launch_bwd_variant<true, false, false>(te, tm, tde, tout, a, (int)grid, st);
    else if (last) rc = x3 ? launch_bwd_variant<false, true, true>(te, tm, tde, tout, a, (int)grid, st)
                           : launch_bwd_variant<false, true, false>(te, tm, tde, tout, a, (int)grid, st);
    else rc = x3 ? launch_bwd_variant<false, false, true>(te, tm, tde, tout, a, (int)grid, st)
                 : launch_bwd_variant<false, false, false>(te, tm, tde, tout, a, (int)grid, st);
    prof_end(tok, st);
    if (rc) return rc;
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

}  // namespace gcrl
