// pna.cu -- principal-neighbourhood aggregation over destination segments.
// Replaces the five torch_scatter launches per GNNBlock.aggregate (architecture.py:62-66,
// principal_neighbourhood_aggregation.py:41-64) by ONE streaming pass over msg[E,F]:
// mean | min | max | std, std = sqrt(relu(mean(x^2) - mean(x)^2) + 1e-5).
//
// HBM-bound: algorithmic bytes = 4*F*E (read) + 16*F*N (write) + 4*(N+1) (seg_off).
// F == 64 kernel: one warp per destination vertex; lanes 0-15 stream the even rows of the
// segment, lanes 16-31 the odd rows, one 128-bit load per lane per row pair (a warp-wide
// request covers two full 256 B rows, fully coalesced), 4 row pairs in flight per warp.
#include "common.cuh"

namespace gcrl {

struct Stat {
    float sum, sq, mn, mx;
    int32_t amn, amx;
};

__device__ __forceinline__ void stat_init(Stat& s) {
    s.sum = 0.f; s.sq = 0.f; s.mn = INFINITY; s.mx = -INFINITY; s.amn = 0x7fffffff; s.amx = 0x7fffffff;
}
// rows must be fed in ascending slot order: strict comparisons keep the FIRST extremum
// (torch_scatter CPU: `if (src < out) {out = src; arg = e}` in edge order).
__device__ __forceinline__ void stat_add(Stat& s, float v, int32_t slot) {
    s.sum += v;
    s.sq = fmaf(v, v, s.sq);
    if (v < s.mn) { s.mn = v; s.amn = slot; }
    if (v > s.mx) { s.mx = v; s.amx = slot; }
}
// merge b into a; ties resolved towards the lower slot id
__device__ __forceinline__ void stat_merge(Stat& a, const Stat& b) {
    a.sum += b.sum;
    a.sq += b.sq;
    if (b.mn < a.mn || (b.mn == a.mn && b.amn < a.amn)) { a.mn = b.mn; a.amn = b.amn; }
    if (b.mx > a.mx || (b.mx == a.mx && b.amx < a.amx)) { a.mx = b.mx; a.amx = b.amx; }
}
// pna.py:45-46,57-64 -- explicit non-contracted mul/sub: var is cancellation-prone and the
// reference evaluates mean_squares - mean*mean with separate roundings.
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

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// ---- F == 64: warp per vertex ------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pna_fwd_w64(const float* __restrict__ msg,
                                                    const int32_t* __restrict__ seg_off, int64_t N,
                                                    float* __restrict__ out,
                                                    int32_t* __restrict__ arg_min,
                                                    int32_t* __restrict__ arg_max,
                                                    float* __restrict__ var_out) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int half = lane >> 4;         // 0: even rows, 1: odd rows
    const int c4 = (lane & 15);         // float4 column
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = warp; v < N; v += nwarps) {
        const int32_t a = seg_off[v], b = seg_off[v + 1];
        Stat s[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) stat_init(s[j]);
        const float4* base = reinterpret_cast<const float4*>(msg) + (int64_t)a * 16 + c4;
        const int32_t deg = b - a;
        int32_t r = half;
        // 4 row pairs (8 rows) in flight per warp
        for (; r + 6 < deg; r += 8) {
            float4 x0 = ldg_stream(base + (int64_t)(r) * 16);
            float4 x1 = ldg_stream(base + (int64_t)(r + 2) * 16);
            float4 x2 = ldg_stream(base + (int64_t)(r + 4) * 16);
            float4 x3 = ldg_stream(base + (int64_t)(r + 6) * 16);
            const float4 xs[4] = {x0, x1, x2, x3};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int32_t slot = a + r + 2 * u;
                stat_add(s[0], xs[u].x, slot);
                stat_add(s[1], xs[u].y, slot);
                stat_add(s[2], xs[u].z, slot);
                stat_add(s[3], xs[u].w, slot);
            }
        }
        for (; r < deg; r += 2) {
            float4 x0 = ldg_stream(base + (int64_t)r * 16);
            int32_t slot = a + r;
            stat_add(s[0], x0.x, slot);
            stat_add(s[1], x0.y, slot);
            stat_add(s[2], x0.z, slot);
            stat_add(s[3], x0.w, slot);
        }
        // combine even-row and odd-row partials
        float res[4][5];
        int32_t ramn[4], ramx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            Stat o;
            o.sum = __shfl_xor_sync(0xffffffffu, s[j].sum, 16);
            o.sq = __shfl_xor_sync(0xffffffffu, s[j].sq, 16);
            o.mn = __shfl_xor_sync(0xffffffffu, s[j].mn, 16);
            o.mx = __shfl_xor_sync(0xffffffffu, s[j].mx, 16);
            o.amn = __shfl_xor_sync(0xffffffffu, s[j].amn, 16);
            o.amx = __shfl_xor_sync(0xffffffffu, s[j].amx, 16);
            Stat m;
            if (half == 0) { m = s[j]; stat_merge(m, o); }   // even + odd, fixed order
            else { m = o; stat_merge(m, s[j]); }
            stat_final(m, deg, res[j][0], res[j][1], res[j][2], res[j][3], res[j][4]);
            ramn[j] = m.amn; ramx[j] = m.amx;
        }
        // half 0 writes mean|min, half 1 writes max|std (both halves hold the full result)
        float4* o4 = reinterpret_cast<float4*>(out + v * 256);
        if (half == 0) {
            o4[c4] = make_float4(res[0][0], res[1][0], res[2][0], res[3][0]);
            o4[16 + c4] = make_float4(res[0][1], res[1][1], res[2][1], res[3][1]);
            if (arg_min) reinterpret_cast<int4*>(arg_min + v * 64)[c4] = make_int4(ramn[0], ramn[1], ramn[2], ramn[3]);
            if (var_out) reinterpret_cast<float4*>(var_out + v * 64)[c4] = make_float4(res[0][4], res[1][4], res[2][4], res[3][4]);
        } else {
            o4[32 + c4] = make_float4(res[0][2], res[1][2], res[2][2], res[3][2]);
            o4[48 + c4] = make_float4(res[0][3], res[1][3], res[2][3], res[3][3]);
            if (arg_max) reinterpret_cast<int4*>(arg_max + v * 64)[c4] = make_int4(ramx[0], ramx[1], ramx[2], ramx[3]);
        }
    }
}

// ---- generic width: thread per (vertex, column) -------------------------------------------
__global__ void k_pna_fwd_generic(const float* __restrict__ msg, const int32_t* __restrict__ seg_off,
                                  int64_t N, int32_t F, float* __restrict__ out,
                                  int32_t* __restrict__ arg_min, int32_t* __restrict__ arg_max,
                                  float* __restrict__ var_out) {
    pdl_prologue();
    int64_t total = N * F;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = t / F;
        int32_t c = (int32_t)(t - v * F);
        int32_t a = seg_off[v], b = seg_off[v + 1];
        Stat s;
        stat_init(s);
        for (int32_t r = a; r < b; ++r) stat_add(s, msg[(int64_t)r * F + c], r);
        float mean, mn, mx, sd, var;
        stat_final(s, b - a, mean, mn, mx, sd, var);
        float* o = out + v * 4 * F;
        o[c] = mean; o[F + c] = mn; o[2 * F + c] = mx; o[3 * F + c] = sd;
        if (arg_min) arg_min[t] = s.amn;
        if (arg_max) arg_max[t] = s.amx;
        if (var_out) var_out[t] = var;
    }
}

__global__ void k_segment_reduce(const float* __restrict__ msg, const int32_t* __restrict__ seg_off,
                                 int64_t N, int32_t F, int which, float* __restrict__ out) {
    pdl_prologue();
    int64_t total = N * F;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = t / F;
        int32_t c = (int32_t)(t - v * F);
        int32_t a = seg_off[v], b = seg_off[v + 1];
        Stat s;
        stat_init(s);
        for (int32_t r = a; r < b; ++r) stat_add(s, msg[(int64_t)r * F + c], r);
        float mean, mn, mx, sd, var;
        stat_final(s, b - a, mean, mn, mx, sd, var);
        float r;
        switch (which) {
            case 0: r = s.sum; break;
            case 1: r = mean; break;
            case 2: r = mn; break;
            case 3: r = mx; break;
            case 4: r = var; break;
            default: r = sd; break;
        }
        out[t] = r;
    }
}

// ---- backward: elementwise over edges -----------------------------------------------------
// d_msg[s,c] = d_mean/deg + [s==argmin] d_min + [s==argmax] d_max
//            + [var>0] d_std * (msg - mean) / (deg * std)
template <int VEC>
__global__ void k_pna_bwd(const float* __restrict__ msg, const float* __restrict__ out,
                          const float* __restrict__ var, const float* __restrict__ d_out,
                          const int32_t* __restrict__ seg_off, const int32_t* __restrict__ dst,
                          const int32_t* __restrict__ arg_min, const int32_t* __restrict__ arg_max,
                          int64_t E, int32_t F, float* __restrict__ d_msg) {
    pdl_prologue();
    int64_t total = E * F;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t s = t / F;
        int32_t c = (int32_t)(t - s * F);
        int32_t v = dst[s];
        float deg = (float)(seg_off[v + 1] - seg_off[v]);
        const float* o = out + (int64_t)v * 4 * F;
        const float* g = d_out + (int64_t)v * 4 * F;
        float m = msg[t];
        float r = g[c] / deg;
        if (arg_min[(int64_t)v * F + c] == (int32_t)s) r += g[F + c];
        if (arg_max[(int64_t)v * F + c] == (int32_t)s) r += g[2 * F + c];
        bool pos = var ? (var[(int64_t)v * F + c] > 0.f) : true;
        if (pos) r += g[3 * F + c] * (m - o[c]) / (deg * o[3 * F + c]);
        d_msg[t] = r;
    }
}

}  // namespace gcrl

using namespace gcrl;

extern "C" {

int gcrl_pna_forward(const float* msg, const int32_t* seg_off, int64_t N, int64_t E,
                        int32_t F, float* out, int32_t* arg_min, int32_t* arg_max,
                        float* var_out, void* stream) {
    GCRL_CHECK_ARG(N >= 0 && E >= 0 && F >= 1);
    if (N == 0) return GCRL_OK;
    GCRL_CHECK_ARG(seg_off && out && (E == 0 || msg));
    cudaStream_t st = (cudaStream_t)stream;
    const int tok = prof_begin(GCRL_PROF_PNA_FWD, st);
    if (F == 64 && (((uintptr_t)msg | (uintptr_t)out) & 15) == 0) {
        // 8 warps per CTA; enough CTAs for >= 2 waves at 8 CTAs/SM, capped (grid-stride)
        int64_t blocks = (N + 7) / 8;
        int64_t cap = (int64_t)sm_count() * 8 * 4;
        if (blocks > cap) blocks = cap;
        launch_k(k_pna_fwd_w64, dim3((int)blocks), dim3(256), 0, st, msg, seg_off, N, out, arg_min, arg_max, var_out);
    } else {
        int64_t total = N * F;
        int64_t blocks = (total + 255) / 256;
        int64_t cap = (int64_t)sm_count() * 32;
        if (blocks > cap) blocks = cap;
        launch_k(k_pna_fwd_generic, dim3((int)blocks), dim3(256), 0, st, msg, seg_off, N, F, out, arg_min, arg_max, var_out);
    }
    prof_end(tok, st);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

int gcrl_pna_backward(const float* msg, const float* out, const float* var, const float* d_out,
                         const int32_t* seg_off, const int32_t* dst, const int32_t* arg_min,
                         const int32_t* arg_max, int64_t N, int64_t E, int32_t F, float* d_msg,
                         void* stream) {
    GCRL_CHECK_ARG(N >= 0 && E >= 0 && F >= 1);
    if (E == 0) return GCRL_OK;
    GCRL_CHECK_ARG(msg && out && var && d_out && seg_off && dst && arg_min && arg_max && d_msg);
    int64_t total = E * F;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    const int tok = prof_begin(GCRL_PROF_PNA_BWD, (cudaStream_t)stream);
    launch_k(k_pna_bwd<1>, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, msg, out, var, d_out, seg_off, dst,
                                                              arg_min, arg_max, E, F, d_msg);
    prof_end(tok, (cudaStream_t)stream);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

int gcrl_segment_reduce(const float* msg, const int32_t* seg_off, int64_t N, int64_t E, int32_t F,
                        int which, float* out, void* stream) {
    GCRL_CHECK_ARG(N >= 0 && E >= 0 && F >= 1 && which >= 0 && which <= 5);
    if (N == 0) return GCRL_OK;
    GCRL_CHECK_ARG(seg_off && out && (E == 0 || msg));
    int64_t total = N * F;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    launch_k(k_segment_reduce, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, msg, seg_off, N, F, which, out);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

}  // extern "C"
