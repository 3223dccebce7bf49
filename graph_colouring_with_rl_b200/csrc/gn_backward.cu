// gn_backward.cu -- backward of the GN Q-network (autograd of dqn_agent.py:219 through
// architecture.py:179-199), fp32 parity path.
//
// Per GNNBlock, walking l = 5..1:
//   vertex side (N rows, minor): small strided SGEMMs -- update-MLP dgrad/wgrad, then after the
//   edge kernel the W1 vertex-column wgrads (dA = dPi^T h, dB = dPj^T h) and
//   dh_{l-1} += dPi A + dPj B.
//   edge side (E rows, dominant): ONE fused kernel per block, k_edge_bwd.  Per 128-slot tile it
//   re-derives hid = relu(Pi[dst] + Pj[src] + C e_{l-1}) from the stashed e_{l-1} (so no
//   hidden activation is ever stored), forms dm = d e_l + PNA-backward(d agg), then
//   dW2 += dm^T hid, d hid = dm W2, dz1 = d hid * relu', dC += dz1^T e_{l-1},
//   d e_{l-1} = dz1 C, dPi[dst] += dz1 (segment partial sums), dPj[src] += dz1 (vector reds).
//   HBM traffic per edge: read e_{l-1}, e_l, d e_l; write d e_{l-1} (1 KB fp32); block 5
//   recomputes e_5 instead of reading it, block 1 writes no d e_0.
#include "gn_common.cuh"
#include <stdlib.h>

namespace gcrl {

// =========================== generic strided SGEMM (vertex side) ===========================
// C(m,n) (op)= sum_k A(m,k) B(k,n); every operand addressed through element strides.
struct GemmArgs {
    const float* A; int64_t sa_m, sa_k;
    const float* B; int64_t sb_k, sb_n;
    float* C;       int64_t sc_m, sc_n;
    int64_t M, N, K;
    const float* bias;   // indexed by n, or null
    const float* mask;   // indexed like C; result multiplied by (mask > 0), or null
    int accumulate;      // C += result
    int relu;            // C = max(C, 0) after everything else
    int atomic;          // atomicAdd result into C (split-K over blockIdx.z)
    int64_t k_chunk;     // K range per blockIdx.z
};

constexpr int GB = 64;   // tile M = N
constexpr int GK = 16;   // tile K

__global__ void __launch_bounds__(256) k_sgemm(const GemmArgs g) {
    pdl_prologue();
    __shared__ __align__(16) float As[GK][GB + 4];
    __shared__ __align__(16) float Bs[GK][GB + 4];
    const int tid = threadIdx.x;
    const int tm = tid >> 4, tn = tid & 15;
    const int64_t m0 = (int64_t)blockIdx.x * GB, n0 = (int64_t)blockIdx.y * GB;
    const int64_t kbeg = (int64_t)blockIdx.z * g.k_chunk;
    const int64_t kend = min(g.K, kbeg + g.k_chunk);
    const bool a_kfast = (g.sa_k == 1);
    const bool b_nfast = (g.sb_n == 1);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int64_t k0 = kbeg; k0 < kend; k0 += GK) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int idx = it * 256 + tid;
            int mm, kk;
            if (a_kfast) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
            const int64_t m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < g.M && k < kend) ? g.A[m * g.sa_m + k * g.sa_k] : 0.f;
            int nn, kb;
            if (b_nfast) { nn = idx & 63; kb = idx >> 6; } else { kb = idx & 15; nn = idx >> 4; }
            const int64_t n = n0 + nn, k2 = k0 + kb;
            Bs[kb][nn] = (n < g.N && k2 < kend) ? g.B[k2 * g.sb_k + n * g.sb_n] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][4 * tm]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][4 * tn]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + 4 * tm + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t n = n0 + 4 * tn + j;
            if (n >= g.N) continue;
            const int64_t ci = m * g.sc_m + n * g.sc_n;
            float v = acc[i][j];
            if (g.bias && blockIdx.z == 0) v += g.bias[n];
            if (g.mask) v = (g.mask[ci] > 0.f) ? v : 0.f;
            if (g.atomic) { atomicAdd(&g.C[ci], v); continue; }
            if (g.accumulate) v += g.C[ci];
            if (g.relu) v = fmaxf(v, 0.f);
            g.C[ci] = v;
        }
    }
}

static int launch_gemm(GemmArgs g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0) return GCRL_OK;
    if (g.K <= 0 && (g.atomic || g.accumulate)) return GCRL_OK;
    if (g.atomic) {
        // split-K so the reduction over vertices fills the machine
        int64_t tiles = ((g.M + GB - 1) / GB) * ((g.N + GB - 1) / GB);
        int64_t want = (int64_t)sm_count() * 4 / (tiles > 0 ? tiles : 1);
        if (want < 1) want = 1;
        int64_t chunk = (g.K + want - 1) / want;
        chunk = (chunk + GK - 1) / GK * GK;
        if (chunk < 64) chunk = 64;          // each CTA's K loop is a chain of dependent global-load rounds: keep it short
        g.k_chunk = chunk;
    } else {
        g.k_chunk = g.K > 0 ? g.K : 1;
    }
    int64_t gz = g.K > 0 ? (g.K + g.k_chunk - 1) / g.k_chunk : 1;
    dim3 grid((unsigned)((g.M + GB - 1) / GB), (unsigned)((g.N + GB - 1) / GB), (unsigned)gz);
    launch_k(k_sgemm, dim3(grid), dim3(256), 0, st, g);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

// dW[o][wcol + k] += sum_r Y[r][o] X[r][k]      (Y [R,ldy], X [R,ldx], dW [O, ldw])
static int wgrad(const float* Y, int ldy, int O, const float* X, int ldx, int K, float* dW, int ldw,
                 int wcol, int64_t R, cudaStream_t st) {
    GemmArgs g{};
    g.A = Y; g.sa_m = 1; g.sa_k = ldy;
    g.B = X; g.sb_k = ldx; g.sb_n = 1;
    g.C = dW + wcol; g.sc_m = ldw; g.sc_n = 1;
    g.M = O; g.N = K; g.K = R;
    g.atomic = 1;
    return launch_gemm(g, st);
}

// dX[r][k] (+)= sum_o dY[r][o] W[o][wcol + k], optionally masked by (mask[r][k] > 0)
static int dgrad(const float* dY, int ldy, int O, const float* W, int ldw, int wcol, int K, float* dX,
                 int ldx, const float* mask, int accumulate, int64_t R, cudaStream_t st) {
    GemmArgs g{};
    g.A = dY; g.sa_m = ldy; g.sa_k = 1;
    g.B = W + wcol; g.sb_k = ldw; g.sb_n = 1;
    g.C = dX; g.sc_m = ldx; g.sc_n = 1;
    g.M = R; g.N = K; g.K = O;
    g.mask = mask; g.accumulate = accumulate;
    return launch_gemm(g, st);
}

// out[c] += sum_r Y[r][c], width <= 64
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ Y, int64_t R, int width,
                                               float* __restrict__ out) {
    pdl_prologue();
    __shared__ float part[4][64];
    const int c = threadIdx.x & 63, q = threadIdx.x >> 6;
    float s = 0.f;
    if (c < width)
        for (int64_t r = (int64_t)blockIdx.x * 4 + q; r < R; r += (int64_t)gridDim.x * 4) s += Y[r * width + c];
    part[q][c] = s;
    __syncthreads();
    if (q == 0 && c < width) atomicAdd(&out[c], (part[0][c] + part[1][c]) + (part[2][c] + part[3][c]));
}

static int colsum(const float* Y, int64_t R, int width, float* out, cudaStream_t st) {
    if (R <= 0) return GCRL_OK;
    int64_t blocks = (R + 3) / 4;
    int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    launch_k(k_colsum, dim3((int)blocks), dim3(256), 0, st, Y, R, width, out);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

__global__ void k_gather_rows(const float* __restrict__ table, const int32_t* __restrict__ idx,
                              int64_t R, int width, float* __restrict__ out) {
    pdl_prologue();
    int64_t total = R * width;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = t / width;
        int c = (int)(t - r * width);
        out[t] = table[(int64_t)idx[r] * width + c];
    }
}

__global__ void k_add_inplace(float* __restrict__ a, const float* __restrict__ b, int64_t n) {
    pdl_prologue();
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
         t += (int64_t)gridDim.x * blockDim.x)
        a[t] += b[t];
}

// ==================================== edge backward ========================================
struct EdgeBwdSmem {
    float Wa[64][64];    // [k][o] = C[o][k]   forward operand of message_mlp.0 (edge part)
    float Wb[64][64];    // [k][o] = W2[o][k]  forward operand of message_mlp.2
    float Cn[64][64];    // [o][k] = C[o][k]   d e_prev = dz1 . C
    float W2n[64][64];   // [o][k] = W2[o][k]  d hid = dm . W2
    float A[TILE][LDT];  // e_{l-1} tile
    float H[TILE][LDT];  // hid, then dz1
    float M[TILE][LDT];  // m = e_l tile, then dm
    float b2[64];
    int32_t dst[TILE];
    int32_t src[TILE];
};

struct EdgeBwdArgs {
    const float* Pi; const float* Pj;
    const float* e_prev; int De;
    const float* W1; int ldW1, coff;
    const float* W2; const float* b2;
    const int32_t* seg_off; const int32_t* src; const int32_t* dst;
    int32_t N, E;
    const float* m;        // stashed e_l [E,64] or null (recompute; last block)
    const float* de;       // d e_l [E,64] or null (last block: e_5 is discarded)
    const float* agg; const float* var; const int32_t* amin; const int32_t* amax;
    const float* dagg;     // [N,256]
    float* de_prev;        // [E,64] or null (first block)
    float* dPi; float* dPj;  // [N,64], pre-zeroed, accumulated with reds
    float* gW1; float* gW2; float* gb2;  // weight gradients (accumulated with reds)
};

// accW[i][j] += sum_r Y[r][4*og+i] * X[r][4*kg+j]
__device__ __forceinline__ void wgrad_tile(const float (*Y)[LDT], const float (*X)[LDT], int og, int kg,
                                           float accW[4][4], float accb[4], bool want_bias) {
#pragma unroll 4
    for (int r = 0; r < TILE; ++r) {
        const float4 y = *reinterpret_cast<const float4*>(&Y[r][4 * og]);
        const float4 x = *reinterpret_cast<const float4*>(&X[r][4 * kg]);
        const float yv[4] = {y.x, y.y, y.z, y.w};
        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) accW[i][j] = fmaf(yv[i], xv[j], accW[i][j]);
            if (want_bias) accb[i] += yv[i];
        }
    }
}

template <bool FIRST>
__global__ void __launch_bounds__(256, 1) k_edge_bwd(const EdgeBwdArgs a) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EdgeBwdSmem& S = *reinterpret_cast<EdgeBwdSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int cg = tid & 15, rg = tid >> 4;   // row-tile mapping: 8 rows x 4 cols per thread
    const int kg = tid & 15, og = tid >> 4;   // wgrad mapping: 4 outputs x 4 inputs per thread

    for (int t = tid; t < 64 * 64; t += 256) {
        const int o = t & 63, k = t >> 6;
        const float w2 = a.W2[o * 64 + k];
        S.Wb[k][o] = w2;
        S.W2n[o][k] = w2;
        if (!FIRST || k < a.De) {
            const float c = a.W1[o * a.ldW1 + a.coff + k];
            S.Wa[k][o] = c;
            S.Cn[o][k] = c;
        }
    }
    if (tid < 64) S.b2[tid] = a.b2[tid];
    __syncthreads();

    float gW2r[4][4], gCr[4][4], gb2r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        gb2r[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { gW2r[i][j] = 0.f; gCr[i][j] = 0.f; }
    }
    float gC_first = 0.f;   // FIRST: thread (o = tid & 63, d = tid >> 6) for d < De

    const int64_t n_tiles = ((int64_t)a.E + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int32_t t0 = (int32_t)(tile * TILE);
        const int nrows = min(TILE, a.E - t0);
        if (tid < TILE) {
            const bool ok = tid < nrows;
            S.dst[tid] = ok ? a.dst[t0 + tid] : 0;
            S.src[tid] = ok ? a.src[t0 + tid] : 0;
        }
        if (!FIRST) {
            const float4* g = reinterpret_cast<const float4*>(a.e_prev + (int64_t)t0 * 64);
            const float4* gm = a.m ? reinterpret_cast<const float4*>(a.m + (int64_t)t0 * 64) : nullptr;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int idx = it * 256 + tid;
                const int r = idx >> 4, c4 = idx & 15;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f), mv = v;
                if (r < nrows) { v = __ldcs(g + idx); if (gm) mv = __ldcs(gm + idx); }
                *reinterpret_cast<float4*>(&S.A[r][4 * c4]) = v;
                if (gm) *reinterpret_cast<float4*>(&S.M[r][4 * c4]) = mv;
            }
        } else {
            for (int t = tid; t < TILE * a.De; t += 256) {
                const int r = t / a.De, d = t - r * a.De;
                S.A[r][d] = (r < nrows) ? a.e_prev[(int64_t)(t0 + r) * a.De + d] : 0.f;
            }
            if (a.m) {
                const float4* gm = reinterpret_cast<const float4*>(a.m + (int64_t)t0 * 64);
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int idx = it * 256 + tid;
                    const int r = idx >> 4, c4 = idx & 15;
                    float4 mv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < nrows) mv = __ldcs(gm + idx);
                    *reinterpret_cast<float4*>(&S.M[r][4 * c4]) = mv;
                }
            }
        }
        __syncthreads();

        // ---- recompute hid = relu(Pi[dst] + Pj[src] + C e_prev) (same op order as forward) ---
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
        if (!FIRST) {
            gemm64(S.A, S.Wa, rg, cg, acc);
        } else {
            for (int d = 0; d < a.De; ++d) {
                const float4 wv = *reinterpret_cast<const float4*>(&S.Wa[d][4 * cg]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float ev = S.A[8 * rg + i][d];
                    acc[i][0] = fmaf(ev, wv.x, acc[i][0]);
                    acc[i][1] = fmaf(ev, wv.y, acc[i][1]);
                    acc[i][2] = fmaf(ev, wv.z, acc[i][2]);
                    acc[i][3] = fmaf(ev, wv.w, acc[i][3]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 8 * rg + i;
            float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrows) {
                const float4 pi = __ldg(reinterpret_cast<const float4*>(a.Pi + (int64_t)S.dst[r] * 64) + cg);
                const float4 pj = __ldg(reinterpret_cast<const float4*>(a.Pj + (int64_t)S.src[r] * 64) + cg);
                hv.x = fmaxf((pi.x + pj.x) + acc[i][0], 0.f);
                hv.y = fmaxf((pi.y + pj.y) + acc[i][1], 0.f);
                hv.z = fmaxf((pi.z + pj.z) + acc[i][2], 0.f);
                hv.w = fmaxf((pi.w + pj.w) + acc[i][3], 0.f);
            }
            *reinterpret_cast<float4*>(&S.H[r][4 * cg]) = hv;
        }
        __syncthreads();

        // ---- m (recomputed for the last block) and dm = d e_l + PNA backward -----------------
        if (!a.m) {
            const float4 bb = *reinterpret_cast<const float4*>(&S.b2[4 * cg]);
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc[i][0] = bb.x; acc[i][1] = bb.y; acc[i][2] = bb.z; acc[i][3] = bb.w; }
            gemm64(S.H, S.Wb, rg, cg, acc);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 mv = *reinterpret_cast<const float4*>(&S.M[8 * rg + i][4 * cg]);
                acc[i][0] = mv.x; acc[i][1] = mv.y; acc[i][2] = mv.z; acc[i][3] = mv.w;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 8 * rg + i;
            float4 dm = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrows) {
                const int32_t v = S.dst[r];
                const int32_t slot = t0 + r;
                const float deg = (float)(a.seg_off[v + 1] - a.seg_off[v]);
                const float4* gp = reinterpret_cast<const float4*>(a.dagg + (int64_t)v * 256);
                const float4* op = reinterpret_cast<const float4*>(a.agg + (int64_t)v * 256);
                const float4 gmean = __ldg(gp + cg), gmin = __ldg(gp + 16 + cg);
                const float4 gmax = __ldg(gp + 32 + cg), gstd = __ldg(gp + 48 + cg);
                const float4 mean = __ldg(op + cg), sd = __ldg(op + 48 + cg);
                const float4 vr = __ldg(reinterpret_cast<const float4*>(a.var + (int64_t)v * 64) + cg);
                const int4 imn = __ldg(reinterpret_cast<const int4*>(a.amin + (int64_t)v * 64) + cg);
                const int4 imx = __ldg(reinterpret_cast<const int4*>(a.amax + (int64_t)v * 64) + cg);
                if (a.de) dm = __ldcs(reinterpret_cast<const float4*>(a.de + (int64_t)slot * 64) + cg);
                dm.x += gmean.x / deg; dm.y += gmean.y / deg; dm.z += gmean.z / deg; dm.w += gmean.w / deg;
                if (imn.x == slot) dm.x += gmin.x;
                if (imn.y == slot) dm.y += gmin.y;
                if (imn.z == slot) dm.z += gmin.z;
                if (imn.w == slot) dm.w += gmin.w;
                if (imx.x == slot) dm.x += gmax.x;
                if (imx.y == slot) dm.y += gmax.y;
                if (imx.z == slot) dm.z += gmax.z;
                if (imx.w == slot) dm.w += gmax.w;
                if (vr.x > 0.f) dm.x += gstd.x * (acc[i][0] - mean.x) / (deg * sd.x);
                if (vr.y > 0.f) dm.y += gstd.y * (acc[i][1] - mean.y) / (deg * sd.y);
                if (vr.z > 0.f) dm.z += gstd.z * (acc[i][2] - mean.z) / (deg * sd.z);
                if (vr.w > 0.f) dm.w += gstd.w * (acc[i][3] - mean.w) / (deg * sd.w);
            }
            *reinterpret_cast<float4*>(&S.M[8 * rg + i][4 * cg]) = dm;   // own elements only
        }
        __syncthreads();

        // ---- dW2 += dm^T hid, db2 += colsum(dm) ----------------------------------------------
        wgrad_tile(S.M, S.H, og, kg, gW2r, gb2r, kg == 0);

        // ---- d hid = dm W2 ; dz1 = d hid * (hid > 0) -----------------------------------------
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
        gemm64(S.M, S.W2n, rg, cg, acc);
        unsigned maskbits = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 hv = *reinterpret_cast<const float4*>(&S.H[8 * rg + i][4 * cg]);
            maskbits |= (hv.x > 0.f ? 1u : 0u) << (4 * i);
            maskbits |= (hv.y > 0.f ? 1u : 0u) << (4 * i + 1);
            maskbits |= (hv.z > 0.f ? 1u : 0u) << (4 * i + 2);
            maskbits |= (hv.w > 0.f ? 1u : 0u) << (4 * i + 3);
        }
        __syncthreads();   // every thread is done reading hid
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 z;
            z.x = (maskbits >> (4 * i)) & 1u ? acc[i][0] : 0.f;
            z.y = (maskbits >> (4 * i + 1)) & 1u ? acc[i][1] : 0.f;
            z.z = (maskbits >> (4 * i + 2)) & 1u ? acc[i][2] : 0.f;
            z.w = (maskbits >> (4 * i + 3)) & 1u ? acc[i][3] : 0.f;
            *reinterpret_cast<float4*>(&S.H[8 * rg + i][4 * cg]) = z;
        }
        __syncthreads();

        // ---- dC += dz1^T e_prev ---------------------------------------------------------------
        if (!FIRST) {
            float dummy[4] = {0.f, 0.f, 0.f, 0.f};
            wgrad_tile(S.H, S.A, og, kg, gCr, dummy, false);
        } else {
            const int o = tid & 63, d = tid >> 6;
            if (d < a.De)
                for (int r = 0; r < TILE; ++r) gC_first = fmaf(S.H[r][o], S.A[r][d], gC_first);
            // De > 4 (never in the reference: edge_features_dim = 1) handled below
            for (int d2 = d + 4; d2 < a.De; d2 += 4) {
                float s = 0.f;
                for (int r = 0; r < TILE; ++r) s = fmaf(S.H[r][o], S.A[r][d2], s);
                atomicAdd(&a.gW1[o * a.ldW1 + a.coff + d2], s);
            }
        }

        // ---- d e_{l-1} = dz1 C -----------------------------------------------------------------
        if (a.de_prev) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
            gemm64(S.H, S.Cn, rg, cg, acc);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = 8 * rg + i;
                if (r < nrows)
                    __stcs(reinterpret_cast<float4*>(a.de_prev + (int64_t)(t0 + r) * 64) + cg,
                           make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
            }
        }

        // ---- dPi[dst] += dz1 (runs of equal dst, one red per run) ------------------------------
        {
            const int col = tid & 63, q = tid >> 6;
            const int r0 = q * 32, r1 = min(r0 + 32, nrows);
            if (r0 < r1) {
                int32_t cd = S.dst[r0];
                float s = 0.f;
                for (int r = r0; r < r1; ++r) {
                    const int32_t d = S.dst[r];
                    if (d != cd) {
                        atomicAdd(&a.dPi[(int64_t)cd * 64 + col], s);
                        s = 0.f;
                        cd = d;
                    }
                    s += S.H[r][col];
                }
                atomicAdd(&a.dPi[(int64_t)cd * 64 + col], s);
            }
        }
        // ---- dPj[src] += dz1 (128-bit vector reds; sources are distinct inside a segment) ------
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int idx = it * 256 + tid;
            const int r = idx >> 4, c4 = idx & 15;
            if (r < nrows) {
                const float4 z = *reinterpret_cast<const float4*>(&S.H[r][4 * c4]);
                atomicAdd(reinterpret_cast<float4*>(a.dPj + (int64_t)S.src[r] * 64) + c4, z);
            }
        }
        __syncthreads();
    }

    // ---- flush the weight-gradient accumulators ------------------------------------------------
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int o = 4 * og + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = 4 * kg + j;
            atomicAdd(&a.gW2[o * 64 + k], gW2r[i][j]);
            if (!FIRST) atomicAdd(&a.gW1[o * a.ldW1 + a.coff + k], gCr[i][j]);
        }
        if (kg == 0) atomicAdd(&a.gb2[o], gb2r[i]);
    }
    if (FIRST) {
        const int o = tid & 63, d = tid >> 6;
        if (d < a.De) atomicAdd(&a.gW1[o * a.ldW1 + a.coff + d], gC_first);
    }
}

static int launch_edge_bwd(const EdgeBwdArgs& a, bool first, cudaStream_t st) {
    if (a.E == 0) return GCRL_OK;
    static bool attr_done = false;
    if (!attr_done) {
        GCRL_CUDA(cudaFuncSetAttribute(k_edge_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EdgeBwdSmem)));
        GCRL_CUDA(cudaFuncSetAttribute(k_edge_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EdgeBwdSmem)));
        attr_done = true;
    }
    int64_t tiles = ((int64_t)a.E + TILE - 1) / TILE;
    int64_t grid = sm_count();
    if (grid > tiles) grid = tiles;
    const int tok = prof_begin(first ? GCRL_PROF_EDGE_BWD_FIRST : (a.m ? GCRL_PROF_EDGE_BWD : GCRL_PROF_EDGE_BWD_LAST), st);
    if (first) launch_k(k_edge_bwd<true>, dim3((int)grid), dim3(256), sizeof(EdgeBwdSmem), st, a);
    else launch_k(k_edge_bwd<false>, dim3((int)grid), dim3(256), sizeof(EdgeBwdSmem), st, a);
    prof_end(tok, st);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

// gn_backward_wg.cu: PNA backward coefficients and the two-group tensor-core edge backward
int launch_pna_bwd_coef(const float* dagg, const float* agg, const float* var, const int32_t* seg_off, int64_t N,
                        float* ca, float* cb, const int32_t* amin, const int32_t* amax, float* de_scatter,
                        float* gb2_vertex, float* zero_ptr, int64_t zero_n, cudaStream_t st);
int launch_edge_bwd_wg(bool first, bool x3, const float* Pi, const float* Pj, const float* e_prev, int De,
                       const float* W1, int ldW1, int coff, const float* W2, const float* b2, const int32_t* seg_off,
                       const int32_t* src, const int32_t* dst, int64_t N, int64_t E, const float* m, const float* de,
                       const float* ca, const float* cb, const float* dagg, const int32_t* amin, const int32_t* amax,
                       float* de_prev, float* dPi, float* dPj, float* gW1, float* gW2, float* gb2, cudaStream_t st,
                       const unsigned char* img_w2, const unsigned char* img_wc, const float* x_first, const float* b1,
                       int Dv);
// gn_node_tc.cu: vertex-side products on the tensor cores
int dgrad64_tc(bool x3, const float* dY, const float* W, int ldw, int wcol, float* dX, int ldx, int xcol, const float* mask,
               int accumulate, int64_t N, cudaStream_t st, const float* dY2 = nullptr, int ngroups = 1);
int wgrad64_tc(bool x3, const float* Y, const float* X1, int ld1, int n1, const float* X2, int ld2, int n2, float* G, int ldG,
               int gcol0, float* gbias, int64_t N, cudaStream_t st);
int linear_bwd_tc(bool x3, const float* Y, const float* X1, int ld1, int n1, const float* X2, int ld2, int n2, float* G, int ldG,
                  int gcol0, float* gbias, const float* W, int ldW, int wcol0, uint32_t dx_chunks, uint32_t mask_chunks,
                  float* dX1, int ldo1, int ocol1, int acc1, float* dX2, int ldo2, int ocol2, int acc2, int64_t N, cudaStream_t st,
                  const unsigned char* wimg);
// GCRL_VBWD_IMPL=split: the vertex side of the backward as separate weight-gradient / data-gradient launches (A/B switch)
static bool vbwd_fused() {
    static const bool split = getenv("GCRL_VBWD_IMPL") && getenv("GCRL_VBWD_IMPL")[0] == 's';
    return !split;
}

// Block 1's narrow weight gradients in one launch (its vertex features are Dv <= 8 wide, so these are [64 x Dv] blocks that the
// tensor-core kernels do not cover): G_m[o][col_m + d] += sum_r Y_m[r][o] x[r][d] for up to three (Y_m, G_m, col_m) and
// gbias[o] += sum_r Y_b[r][o] for Y_b = Y of matrix `bias_m` -- instead of three split-K SGEMM launches and a column sum.
struct NarrowWgradArgs {
    const float* Y[3]; float* G[3]; int ldG[3], col[3];
    int nmat, bias_m;
    float* gbias;
    const float* x; int Dv;
    int64_t N;
};
__global__ void __launch_bounds__(192) k_narrow_wgrad(const NarrowWgradArgs a) {
    pdl_prologue();
    const int m = threadIdx.x >> 6, o = threadIdx.x & 63;
    if (m >= a.nmat) return;
    float acc[8], bs = 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] = 0.f;
    const float* Y = a.Y[m];
    for (int64_t r = blockIdx.x; r < a.N; r += gridDim.x) {
        const float y = Y[r * 64 + o];
        bs += y;
#pragma unroll
        for (int d = 0; d < 8; ++d)
            if (d < a.Dv) acc[d] = fmaf(y, a.x[r * a.Dv + d], acc[d]);
    }
    float* g = a.G[m] + (int64_t)o * a.ldG[m] + a.col[m];
    for (int d = 0; d < a.Dv; ++d) atomicAdd(g + d, acc[d]);
    if (a.gbias && m == a.bias_m) atomicAdd(a.gbias + o, bs);
}

// first stage of the head's backward (architecture.py:197 fc3 and the ReLU in front of it) in one launch:
//   dy2[r][k] = dq[r] F3[k] [y2[r][k] > 0],   gF3[k] += sum_r dq[r] y2[r][k],   gf3 += sum_r dq[r]
__global__ void __launch_bounds__(256) k_head_bwd_first(const float* __restrict__ dq, const float* __restrict__ y2,
                                                       const float* __restrict__ F3, int64_t N, float* __restrict__ dy2,
                                                       float* __restrict__ gF3, float* __restrict__ gf3) {
    pdl_prologue();
    __shared__ float part[4][65];
    const int c = threadIdx.x & 63, q = threadIdx.x >> 6;
    const float w = F3[c];
    float s = 0.f, sb = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * 4 + q; r < N; r += (int64_t)gridDim.x * 4) {
        const float g = dq[r], y = y2[r * 64 + c];
        dy2[r * 64 + c] = y > 0.f ? g * w : 0.f;
        s = fmaf(g, y, s);
        sb += g;
    }
    part[q][c] = s;
    if (c == 0) part[q][64] = sb;
    __syncthreads();
    if (q == 0) {
        atomicAdd(&gF3[c], (part[0][c] + part[1][c]) + (part[2][c] + part[3][c]));
        if (c == 0) atomicAdd(gf3, (part[0][64] + part[1][64]) + (part[2][64] + part[3][64]));
    }
}


}  // namespace gcrl

using namespace gcrl;

#define RC(call) do { int rc_ = (call); if (rc_) return rc_; } while (0)

extern "C" {

size_t gcrl_gn_backward_workspace_bytes(int64_t N, int64_t E) {
    const size_t n64 = align_up((size_t)N * 64 * 4);
    size_t b = 2 * align_up((size_t)E * 64 * 4);   // d e ping-pong
    b += 8 * n64;                                   // dh x2, du1, dPi, dPj, gather scratch, PNA coefficients x2
    b += align_up((size_t)N * 256 * 4);             // d agg
    b += align_up(wimg_bytes());                    // pre-converted weight tiles (tensor-core math modes)
    return b + 256;
}

int gcrl_gn_backward(const float* params, int vdim, int edim, int gdim, int64_t N, int64_t E,
                     const int32_t* seg_off, const int32_t* src, const int32_t* dst, const float* x,
                     const float* eattr, const float* gfeat, const int32_t* vgraph,
                     const void* stash_p, const float* d_q, const float* d_h, float* grad,
                     int accumulate, void* ws, size_t ws_bytes, int math_mode, void* stream) {
    GCRL_CHECK_ARG(vdim >= 1 && vdim <= 64 && edim >= 1 && edim <= 64 && gdim >= 0 && gdim <= 64);
    GCRL_CHECK_ARG(N >= 0 && E >= 0 && N < (1ll << 31) && E < (1ll << 31));
    GCRL_CHECK_ARG(params && grad);
    GCRL_CHECK_ARG(math_mode == GCRL_MATH_FP32 || math_mode == GCRL_MATH_BF16 || math_mode == GCRL_MATH_BF16X3);
    cudaStream_t st = (cudaStream_t)stream;
    NetW net = make_net(params, vdim, edim, gdim);
    if (!accumulate) GCRL_CUDA(cudaMemsetAsync(grad, 0, (size_t)net.total * 4, st));
    if (N == 0) return GCRL_OK;
    GCRL_CHECK_ARG(seg_off && x && stash_p && (d_q || d_h) && (E == 0 || (src && dst && eattr)));
    GCRL_CHECK_ARG(gdim == 0 || (gfeat && vgraph));
    if (!ws || ws_bytes < gcrl_gn_backward_workspace_bytes(N, E)) {
        set_error("gcrl_gn_backward: workspace too small");
        return GCRL_ERR_WORKSPACE;
    }
    NvtxRange nv_all("gcrl_gn_backward");
    Stash S = carve_stash(const_cast<void*>(stash_p), N, E);
    Arena ar(ws, ws_bytes);
    float* deb[2] = {ar.take<float>((size_t)E * 64), ar.take<float>((size_t)E * 64)};
    float* dh = ar.take<float>((size_t)N * 64);
    float* dh_prev = ar.take<float>((size_t)N * 64);
    float* du1 = ar.take<float>((size_t)N * 64);
    float* dPi = ar.take<float>((size_t)N * 64);
    float* dPj = ar.take<float>((size_t)N * 64);
    float* scratch = ar.take<float>((size_t)N * 64);
    float* coef_a = ar.take<float>((size_t)N * 64);
    float* coef_b = ar.take<float>((size_t)N * 64);
    float* dagg = ar.take<float>((size_t)N * 256);
    unsigned char* wimg = ar.take<unsigned char>(wimg_bytes());
    {
        static const bool no_wimg = getenv("GCRL_WIMG") && getenv("GCRL_WIMG")[0] == '0';
        if (math_mode == GCRL_MATH_FP32 || gdim != 0 || no_wimg) wimg = nullptr;
        else RC(launch_build_weight_image(net, wimg, st));
    }
    const unsigned char* img_f1 = wimg ? wimg + (size_t)WIMG_F1 * WIMG_TILE_BYTES : nullptr;
    const unsigned char* img_f2 = wimg ? wimg + (size_t)WIMG_F2 * WIMG_TILE_BYTES : nullptr;
    // gradient views, laid out like params
    auto G = [&](const float* p) { return grad + (p - params); };

    // ---- head (architecture.py:191-197) ---------------------------------------------------------
    const HeadW& hd = net.head;
    const int ldF1 = 64 + hd.Dg;
    const float* h5 = S.h[GCRL_N_BLOCKS - 1];
    {
    NvtxRange nv_h("bwd head");
    if (d_q) {
        float* dy2 = du1;       // reuse
        float* dy1 = scratch;
        static const bool node_ffma_h = getenv("GCRL_NODE_IMPL") && getenv("GCRL_NODE_IMPL")[0] == 'f';
        if (math_mode != GCRL_MATH_FP32 && !node_ffma_h && hd.Dg == 0) {
            // tensor-core modes: the 64-wide products of the head on tcgen05 like the blocks' vertex side, bias gradients
            // fused into the weight-gradient kernels, fc3's three tiny products in one launch (12 launches -> 5)
            const bool x3h = math_mode == GCRL_MATH_BF16X3;
            int64_t blocks = (N + 3) / 4;
            if (blocks > (int64_t)sm_count() * 4) blocks = (int64_t)sm_count() * 4;
            launch_k(k_head_bwd_first, dim3((int)blocks), dim3(256), 0, st, d_q, S.y2, hd.F3, N, dy2, G(hd.F3), G(hd.f3));
            GCRL_LAUNCH_CHECK();
            if (vbwd_fused()) {
                // one launch per Linear layer: weight gradient, bias gradient and (masked) data gradient together
                RC(linear_bwd_tc(x3h, dy2, S.y1, 64, 1, nullptr, 0, 0, G(hd.F2), 64, 0, G(hd.f2), hd.F2, 64, 0, 1u, 1u,
                                 dy1, 64, 0, 0, nullptr, 0, 0, 0, N, st, img_f2));
                RC(linear_bwd_tc(x3h, dy1, h5, 64, 1, nullptr, 0, 0, G(hd.F1), 64, 0, G(hd.f1), hd.F1, 64, 0, 1u, 0u,
                                 dh, 64, 0, 0, nullptr, 0, 0, 0, N, st, img_f1));
            } else {
            RC(wgrad64_tc(x3h, dy2, S.y1, 64, 1, nullptr, 0, 0, G(hd.F2), 64, 0, G(hd.f2), N, st));
            RC(dgrad64_tc(x3h, dy2, hd.F2, 64, 0, dy1, 64, 0, S.y1, 0, N, st));
            RC(wgrad64_tc(x3h, dy1, h5, 64, 1, nullptr, 0, 0, G(hd.F1), 64, 0, G(hd.f1), N, st));
            RC(dgrad64_tc(x3h, dy1, hd.F1, 64, 0, dh, 64, 0, nullptr, 0, N, st));
            }
        } else {
        RC(wgrad(d_q, 1, 1, S.y2, 64, 64, G(hd.F3), 64, 0, N, st));              // dF3
        RC(colsum(d_q, N, 1, G(hd.f3), st));                                     // df3
        RC(dgrad(d_q, 1, 1, hd.F3, 64, 0, 64, dy2, 64, S.y2, 0, N, st));         // dy2 = dq F3 * relu'
        RC(wgrad(dy2, 64, 64, S.y1, 64, 64, G(hd.F2), 64, 0, N, st));            // dF2
        RC(colsum(dy2, N, 64, G(hd.f2), st));
        RC(dgrad(dy2, 64, 64, hd.F2, 64, 0, 64, dy1, 64, S.y1, 0, N, st));       // dy1
        RC(wgrad(dy1, 64, 64, h5, 64, 64, G(hd.F1), ldF1, 0, N, st));            // dF1[:, :64]
        if (hd.Dg > 0) {
            // gathered global features reuse dh_prev as [N, Dg] scratch (Dg <= 64)
            int64_t total = N * hd.Dg;
            int blocks = (int)((total + 255) / 256);
            if (blocks > sm_count() * 16) blocks = sm_count() * 16;
            launch_k(k_gather_rows, dim3(blocks), dim3(256), 0, st, gfeat, vgraph, N, hd.Dg, dh_prev);
            GCRL_LAUNCH_CHECK();
            RC(wgrad(dy1, 64, 64, dh_prev, hd.Dg, hd.Dg, G(hd.F1), ldF1, 64, N, st));
        }
        RC(colsum(dy1, N, 64, G(hd.f1), st));
        RC(dgrad(dy1, 64, 64, hd.F1, ldF1, 0, 64, dh, 64, nullptr, 0, N, st));   // dh5
        }
        if (d_h) {
            int64_t n = N * 64;
            int blocks = (int)((n + 255) / 256);
            if (blocks > sm_count() * 16) blocks = sm_count() * 16;
            launch_k(k_add_inplace, dim3(blocks), dim3(256), 0, st, dh, d_h, n);
            GCRL_LAUNCH_CHECK();
        }
    } else {
        GCRL_CUDA(cudaMemcpyAsync(dh, d_h, (size_t)N * 64 * 4, cudaMemcpyDeviceToDevice, st));
    }

    }
    // ---- blocks 5..1 ------------------------------------------------------------------------------
    const float* de_cur = nullptr;   // d e_l
    for (int l = GCRL_N_BLOCKS - 1; l >= 0; --l) {
        NvtxRange nv_blk("bwd block %d", l + 1);
        const BlockW& b = net.blk[l];
        const float* h_prev = (l == 0) ? x : S.h[l - 1];
        const int ldU1 = 256 + b.Dv;
        const int ldW1 = 2 * b.Dv + b.De;
        // update MLP (architecture.py:68-78); tensor-core math modes run the 64-wide products on tcgen05 (gn_node_tc.cu)
        static const bool node_ffma = getenv("GCRL_NODE_IMPL") && getenv("GCRL_NODE_IMPL")[0] == 'f';
        const bool vtc = math_mode != GCRL_MATH_FP32 && !node_ffma;
        const bool x3 = math_mode == GCRL_MATH_BF16X3;
        if (vtc && vbwd_fused()) {
            // update_mlp.2 and update_mlp.0, one launch each: weight + bias gradient, and the data gradient masked by the
            // stashed relu output (du1) / written to d agg's four windows and dh_{l-1}
            RC(linear_bwd_tc(x3, dh, S.u1[l], 64, 1, nullptr, 0, 0, G(b.U2), 64, 0, G(b.c2), b.U2, 64, 0, 1u, 1u,
                             du1, 64, 0, 0, nullptr, 0, 0, 0, N, st, wimg_tile(wimg, l, WT_U2)));
            const int n2 = b.Dv == 64 ? 1 : 0;
            const uint32_t dx = 0xFu | ((n2 && l > 0) ? 0x10u : 0u);
            RC(linear_bwd_tc(x3, du1, S.agg[l], 256, 4, h_prev, 64, n2, G(b.U1), ldU1, 0, G(b.c1), b.U1, ldU1, 0, dx, 0u,
                             dagg, 256, 0, 0, dh_prev, 64, 0, 0, N, st, wimg_tile(wimg, l, WT_U1)));
            // (block 1's narrow tail of U1 joins the dPi / dPj ones in k_narrow_wgrad after the edge kernel)
        } else if (vtc) {
            // bias gradients ride in the weight-gradient launches; d agg's four 64-column windows are one launch
            RC(wgrad64_tc(x3, dh, S.u1[l], 64, 1, nullptr, 0, 0, G(b.U2), 64, 0, G(b.c2), N, st));
            RC(dgrad64_tc(x3, dh, b.U2, 64, 0, du1, 64, 0, S.u1[l], 0, N, st));
            RC(wgrad64_tc(x3, du1, S.agg[l], 256, 4, h_prev, 64, b.Dv == 64 ? 1 : 0, G(b.U1), ldU1, 0, G(b.c1), N, st));
            if (b.Dv != 64) RC(wgrad(du1, 64, 64, h_prev, b.Dv, b.Dv, G(b.U1), ldU1, 256, N, st));
            RC(dgrad64_tc(x3, du1, b.U1, ldU1, 0, dagg, 256, 0, nullptr, 0, N, st, nullptr, 4));
            if (l > 0) RC(dgrad64_tc(x3, du1, b.U1, ldU1, 256, dh_prev, 64, 0, nullptr, 0, N, st));
        } else {
        RC(wgrad(dh, 64, 64, S.u1[l], 64, 64, G(b.U2), 64, 0, N, st));
        RC(colsum(dh, N, 64, G(b.c2), st));
        RC(dgrad(dh, 64, 64, b.U2, 64, 0, 64, du1, 64, S.u1[l], 0, N, st));
        RC(wgrad(du1, 64, 64, S.agg[l], 256, 256, G(b.U1), ldU1, 0, N, st));
        RC(wgrad(du1, 64, 64, h_prev, b.Dv, b.Dv, G(b.U1), ldU1, 256, N, st));
        RC(colsum(du1, N, 64, G(b.c1), st));
        RC(dgrad(du1, 64, 64, b.U1, ldU1, 0, 256, dagg, 256, nullptr, 0, N, st));
        if (l > 0) RC(dgrad(du1, 64, 64, b.U1, ldU1, 256, 64, dh_prev, 64, nullptr, 0, N, st));
        }
        // message MLP + aggregation (architecture.py:51-66)
        const int64_t dp_floats = (int64_t)(dPj - dPi) + N * 64;                                   // dPi and dPj are adjacent
        const bool tc_edge = E > 0 && math_mode != GCRL_MATH_FP32 && ((l == 0) ? (b.De <= 8) : (b.De == 64));
        if (!tc_edge) GCRL_CUDA(cudaMemsetAsync(dPi, 0, (size_t)dp_floats * 4, st));              // (else cleared by k_pna_bwd_coef)
        float* de_prev = (l > 0) ? deb[l & 1] : nullptr;
        const bool tc_ok = (l == 0) ? (b.De <= 8) : (b.De == 64);
        if (E > 0 && math_mode != GCRL_MATH_FP32 && tc_ok) {
            // the arg-min/max terms are pre-added to d e_l by k_pna_bwd_coef (a workspace buffer nobody else reads: two
            // scattered reds per (vertex, column) instead of two compares per element); block 5 has no d e_l and applies
            // them in its dm loop.  The forward stashed e_5, so block 5 runs as a regular block without the d e_l term.
            float* de_scatter = de_cur ? const_cast<float*>(de_cur) : nullptr;
            RC(launch_pna_bwd_coef(dagg, S.agg[l], S.var[l], seg_off, N, coef_a, coef_b, S.amin[l], S.amax[l], de_scatter,
                                   nullptr, dPi, dp_floats, st));
            RC(launch_edge_bwd_wg(l == 0, math_mode == GCRL_MATH_BF16X3, S.Pi[l], S.Pj[l], (l == 0) ? eattr : S.e[l - 1],
                                  b.De, b.W1, ldW1, 2 * b.Dv, b.W2, b.b2, seg_off, src, dst, N, E, S.e[l], de_cur, coef_a, coef_b,
                                  dagg, S.amin[l], S.amax[l], de_prev, dPi, dPj, G(b.W1), G(b.W2), G(b.b2), st,
                                  wimg_tile(wimg, l, WT_W2), l == 0 ? nullptr : wimg_tile(wimg, l, WT_WC),
                                  l == 0 ? x : nullptr, b.b1, b.Dv));
        } else if (E > 0) {
            EdgeBwdArgs a;
            a.Pi = S.Pi[l]; a.Pj = S.Pj[l];
            a.e_prev = (l == 0) ? eattr : S.e[l - 1]; a.De = b.De;
            a.W1 = b.W1; a.ldW1 = ldW1; a.coff = 2 * b.Dv;
            a.W2 = b.W2; a.b2 = b.b2;
            a.seg_off = seg_off; a.src = src; a.dst = dst; a.N = (int32_t)N; a.E = (int32_t)E;
            a.m = (l == GCRL_N_BLOCKS - 1) ? nullptr : S.e[l];
            a.de = de_cur;
            a.agg = S.agg[l]; a.var = S.var[l]; a.amin = S.amin[l]; a.amax = S.amax[l];
            a.dagg = dagg;
            a.de_prev = de_prev;
            a.dPi = dPi; a.dPj = dPj;
            a.gW1 = G(b.W1); a.gW2 = G(b.W2); a.gb2 = G(b.b2);
            RC(launch_edge_bwd(a, l == 0, st));
        }
        // vertex columns of message_mlp.0: Pi = A h + b1, Pj = B h
        const bool fused_w1 = vtc && b.Dv == 64 && vbwd_fused();
        if (fused_w1) {
            // Pi = A h + b1 and Pj = B h: gradient of the A / B columns of message_mlp.0 and dh_{l-1} += dPi A, += dPj B
            const uint32_t dx = l > 0 ? 1u : 0u;
            RC(linear_bwd_tc(x3, dPi, h_prev, 64, 1, nullptr, 0, 0, G(b.W1), ldW1, 0, G(b.b1), b.W1, ldW1, 0, dx, 0u,
                             dh_prev, 64, 0, 1, nullptr, 0, 0, 0, N, st, wimg_tile(wimg, l, WT_WA)));
            RC(linear_bwd_tc(x3, dPj, h_prev, 64, 1, nullptr, 0, 0, G(b.W1), ldW1, 64, nullptr, b.W1, ldW1, 64, dx, 0u,
                             dh_prev, 64, 0, 1, nullptr, 0, 0, 0, N, st, wimg_tile(wimg, l, WT_WB)));
        } else if (vtc && b.Dv == 64) {
            RC(wgrad64_tc(x3, dPi, h_prev, 64, 1, nullptr, 0, 0, G(b.W1), ldW1, 0, G(b.b1), N, st));
            RC(wgrad64_tc(x3, dPj, h_prev, 64, 1, nullptr, 0, 0, G(b.W1), ldW1, 64, nullptr, N, st));
        } else if (vtc && vbwd_fused() && b.Dv <= 8) {
            NarrowWgradArgs na;
            na.Y[0] = dPi; na.G[0] = G(b.W1); na.ldG[0] = ldW1; na.col[0] = 0;
            na.Y[1] = dPj; na.G[1] = G(b.W1); na.ldG[1] = ldW1; na.col[1] = b.Dv;
            na.Y[2] = du1; na.G[2] = G(b.U1); na.ldG[2] = ldU1; na.col[2] = 256;
            na.nmat = 3; na.bias_m = 0; na.gbias = G(b.b1); na.x = h_prev; na.Dv = b.Dv; na.N = N;
            int64_t blocks = N < (int64_t)sm_count() * 4 ? N : (int64_t)sm_count() * 4;
            launch_k(k_narrow_wgrad, dim3((int)blocks), dim3(192), 0, st, na);
            GCRL_LAUNCH_CHECK();
        } else {
            if (vtc && vbwd_fused() && b.Dv != 64) RC(wgrad(du1, 64, 64, h_prev, b.Dv, b.Dv, G(b.U1), ldU1, 256, N, st));
            RC(wgrad(dPi, 64, 64, h_prev, b.Dv, b.Dv, G(b.W1), ldW1, 0, N, st));
            RC(wgrad(dPj, 64, 64, h_prev, b.Dv, b.Dv, G(b.W1), ldW1, b.Dv, N, st));
            RC(colsum(dPi, N, 64, G(b.b1), st));
        }
        if (l > 0 && fused_w1) {
            float* t = dh; dh = dh_prev; dh_prev = t;
        } else if (l > 0 && vtc) {
            // dh_{l-1} += dPi A + dPj B: two K chunks of one launch
            RC(dgrad64_tc(x3, dPi, b.W1, ldW1, 0, dh_prev, 64, 0, nullptr, 1, N, st, dPj));
            float* t = dh; dh = dh_prev; dh_prev = t;
        } else if (l > 0) {
            RC(dgrad(dPi, 64, 64, b.W1, ldW1, 0, 64, dh_prev, 64, nullptr, 1, N, st));
            RC(dgrad(dPj, 64, 64, b.W1, ldW1, 64, 64, dh_prev, 64, nullptr, 1, N, st));
            float* t = dh; dh = dh_prev; dh_prev = t;
        }
        de_cur = de_prev;
    }
    return GCRL_OK;
}

}  // extern "C"
