// dqn.cu -- the scalar side of the DQN agent on the device: masked per-graph argmax
// (dqn_agent.py:131-134), TD target + MSE loss + dQ (dqn_agent.py:178-215), fused Adam +
// soft target update (architecture.py:173, dqn_agent.py:220-229).  All HBM/latency-bound.
#include "common.cuh"

namespace gcrl {

// one warp per graph; first maximum (lowest vertex index) like np.argmax over the ascending
// unassigned list
__device__ __forceinline__ void warp_masked_argmax(const float* __restrict__ q,
                                                   const int32_t* __restrict__ colour, int32_t v0,
                                                   int32_t n, int lane, float& best, int32_t& arg) {
    best = -INFINITY;
    arg = -1;
    for (int32_t i = lane; i < n; i += 32) {
        if (colour[v0 + i] < 0) {
            const float v = q[v0 + i];
            if (arg < 0 || v > best) { best = v; arg = i; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int32_t oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (oa >= 0 && (arg < 0 || ob > best || (ob == best && oa < arg))) { best = ob; arg = oa; }
    }
}

__global__ void k_score_argmax(const float* __restrict__ q, const int32_t* __restrict__ colour,
                               const int32_t* __restrict__ node_off, int32_t G,
                               int32_t* __restrict__ action, float* __restrict__ q_best) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = warp; g < G; g += nwarps) {
        const int32_t v0 = node_off[g], n = node_off[g + 1] - v0;
        float best;
        int32_t arg;
        warp_masked_argmax(q, colour, v0, n, lane, best, arg);
        if (lane == 0) {
            action[g] = arg;
            if (q_best) q_best[g] = best;
        }
    }
}

__global__ void k_td_target_loss(const float* __restrict__ q_cur, const float* __restrict__ q_next,
                                 const int32_t* __restrict__ next_colour,
                                 const int32_t* __restrict__ node_off, int32_t G,
                                 const int32_t* __restrict__ action, const float* __restrict__ reward,
                                 const int32_t* __restrict__ done, float gamma, float inv_b,
                                 float* __restrict__ q_expected, float* __restrict__ q_target,
                                 float* __restrict__ d_q, float* __restrict__ loss, int32_t* __restrict__ err) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = warp; g < G; g += nwarps) {
        const int32_t v0 = node_off[g], n = node_off[g + 1] - v0;
        float best;
        int32_t arg;
        warp_masked_argmax(q_next, next_colour, v0, n, lane, best, arg);   // :192-203
        for (int32_t i = lane; i < n; i += 32) d_q[v0 + i] = 0.f;
        __syncwarp();
        if (lane == 0) {
            const int32_t act = action[g];
            if (act < 0 || act >= n) {            // a bad / sentinel action: no contribution, no out-of-bounds access
                if (err) atomicCAS(err, 0, (int32_t)(g + 1));
                if (q_expected) q_expected[g] = NAN;
                if (q_target) q_target[g] = NAN;
                continue;
            }
            const int32_t dn = done[g];
            float qn = (dn == 1) ? 0.f : best;                              // :207
            const float tgt = reward[g] + (gamma * qn) * (float)(1 - dn);   // :212
            const float qe = q_cur[v0 + act];                         // :178-184
            const float diff = qe - tgt;
            if (q_expected) q_expected[g] = qe;
            if (q_target) q_target[g] = tgt;
            d_q[v0 + act] = 2.f * diff * inv_b;                       // d mse / d q
            atomicAdd(loss, diff * diff * inv_b);                           // :215 (mean)
        }
    }
}

// torch.optim.Adam single-tensor math (lr, betas, eps; no weight decay / amsgrad):
//   m <- m + (g - m)(1 - b1);  v <- b2 v + (1 - b2) g g
//   p <- p - step_size * m / (sqrt(v) / sqrt(bc2) + eps),  step_size = lr / bc1
// followed by dqn_agent.py:226-229  t <- tau p + (1 - tau) t  on the updated p.
__global__ void k_adam_soft_update(float* __restrict__ p, const float* __restrict__ g,
                                   float* __restrict__ m, float* __restrict__ v,
                                   float* __restrict__ t, int64_t n, float one_minus_b1, float b2,
                                   float one_minus_b2, float step_size, float bc2_sqrt, float eps,
                                   float tau, float one_minus_tau) {
    pdl_prologue();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        float mi = m[i], vi = v[i];
        mi = __fadd_rn(mi, __fmul_rn(__fsub_rn(gi, mi), one_minus_b1));
        vi = __fadd_rn(__fmul_rn(vi, b2), __fmul_rn(__fmul_rn(gi, gi), one_minus_b2));
        const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), eps);
        const float pi = __fadd_rn(p[i], __fmul_rn(-step_size, __fdiv_rn(mi, denom)));
        m[i] = mi;
        v[i] = vi;
        p[i] = pi;
        if (t) t[i] = __fadd_rn(__fmul_rn(tau, pi), __fmul_rn(one_minus_tau, t[i]));
    }
}

__global__ void k_soft_update(const float* __restrict__ b, float* __restrict__ t, int64_t n, float tau,
                              float one_minus_tau) {
    pdl_prologue();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        t[i] = __fadd_rn(__fmul_rn(tau, b[i]), __fmul_rn(one_minus_tau, t[i]));
}

}  // namespace gcrl

using namespace gcrl;

static int warp_grid(int64_t G) {
    int64_t blocks = (G * 32 + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

static int elem_grid(int64_t n) {
    int64_t blocks = (n + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

extern "C" {

int gcrl_score_argmax(const float* q, const int32_t* colour, const int32_t* node_off, int32_t G,
                      int32_t* action, float* q_best, void* stream) {
    GCRL_CHECK_ARG(G >= 0);
    if (G == 0) return GCRL_OK;
    GCRL_CHECK_ARG(q && colour && node_off && action);
    launch_k(k_score_argmax, dim3(warp_grid(G)), dim3(256), 0, (cudaStream_t)stream, q, colour, node_off, G, action, q_best);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

int gcrl_td_target_loss(const float* q_cur, const float* q_next, const int32_t* next_colour,
                        const int32_t* node_off, int32_t G, const int32_t* action,
                        const float* reward, const int32_t* done, float gamma, int32_t batch_total,
                        float* q_expected, float* q_target, float* d_q, float* loss, int32_t* err, void* stream) {
    GCRL_CHECK_ARG(G >= 0 && batch_total >= 1);
    if (G == 0) return GCRL_OK;
    GCRL_CHECK_ARG(q_cur && q_next && next_colour && node_off && action && reward && done && d_q && loss);
    launch_k(k_td_target_loss, dim3(warp_grid(G)), dim3(256), 0, (cudaStream_t)stream, q_cur, q_next, next_colour, node_off, G, action, reward, done, gamma, 1.0f / (float)batch_total,
        q_expected, q_target, d_q, loss, err);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

int gcrl_adam_soft_update(float* params, const float* grad, float* exp_avg, float* exp_avg_sq,
                          float* target, int64_t n, float lr, float beta1, float beta2, float eps,
                          int64_t step, float tau, void* stream) {
    GCRL_CHECK_ARG(n >= 0 && step >= 1);
    if (n == 0) return GCRL_OK;
    GCRL_CHECK_ARG(params && grad && exp_avg && exp_avg_sq);
    // scalar prologue in double like torch's Python-side arithmetic
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    launch_k(k_adam_soft_update, dim3(elem_grid(n)), dim3(256), 0, (cudaStream_t)stream, params, grad, exp_avg, exp_avg_sq, target, n, (float)(1.0 - (double)beta1), beta2,
        (float)(1.0 - (double)beta2), step_size, bc2_sqrt, eps, tau, (float)(1.0 - (double)tau));
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

int gcrl_soft_update(const float* behaviour, float* target, int64_t n, float tau, void* stream) {
    GCRL_CHECK_ARG(n >= 0);
    if (n == 0) return GCRL_OK;
    GCRL_CHECK_ARG(behaviour && target);
    launch_k(k_soft_update, dim3(elem_grid(n)), dim3(256), 0, (cudaStream_t)stream, behaviour, target, n, tau,
                                                                 (float)(1.0 - (double)tau));
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

}  // extern "C"
