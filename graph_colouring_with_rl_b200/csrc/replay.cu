// replay.cu -- device-resident replay ring (dqn_agent.py:8-65 ReplayBuffer).
// The reference keeps 1e5 transitions as pairs of PyG `Data` objects on the host and collates
// 128 of them per learn().  Every graph is complete with static edges, so a transition is fully
// described by (graph id, colours before, colours after, action, reward, done): the ring stores
// exactly that (colours as int16, stride n_max) and the minibatch is materialised by ONE gather
// launch straight into the flat arrays gcrl_pack_dense / gcrl_td_target_loss consume.
// Byte/integer work, HBM-bound: 2*(n^2 + 6n) bytes per gathered transition.
#include "common.cuh"

namespace gcrl {

// one warp per environment: colour vectors -> ring slot
__global__ void k_replay_store(int32_t n_envs, const int32_t* __restrict__ slot,
                               const int32_t* __restrict__ graph_id,
                               const int32_t* __restrict__ node_off,
                               const int32_t* __restrict__ cur_colour,
                               const int32_t* __restrict__ next_colour,
                               const int32_t* __restrict__ action, const float* __restrict__ reward,
                               const int32_t* __restrict__ done, int32_t stride,
                               int32_t* __restrict__ r_graph, int16_t* __restrict__ r_cur,
                               int16_t* __restrict__ r_next, int32_t* __restrict__ r_action,
                               float* __restrict__ r_reward, int32_t* __restrict__ r_done) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = warp; g < n_envs; g += nwarps) {
        const int32_t s = slot[g];
        if (s < 0) continue;
        const int32_t v0 = node_off[g], n = node_off[g + 1] - v0;
        int16_t* c = r_cur + (int64_t)s * stride;
        int16_t* d = r_next + (int64_t)s * stride;
        for (int32_t i = lane; i < stride; i += 32) {
            c[i] = i < n ? (int16_t)cur_colour[v0 + i] : (int16_t)-1;
            d[i] = i < n ? (int16_t)next_colour[v0 + i] : (int16_t)-1;
        }
        if (lane == 0) {
            r_graph[s] = graph_id[g];
            r_action[s] = action[g];
            r_reward[s] = reward[g];
            r_done[s] = done[g];
        }
    }
}

constexpr int GATHER_CHUNK = 16384;   // adjacency bytes per CTA

// grid (chunks, batch): adjacency bytes of the slot's graph -> minibatch position b; chunk 0 also
// widens the colour vectors and copies the scalars
__global__ void k_replay_gather(const int32_t* __restrict__ slots, const int32_t* __restrict__ r_graph,
                                const int16_t* __restrict__ r_cur, const int16_t* __restrict__ r_next,
                                const int32_t* __restrict__ r_action, const float* __restrict__ r_reward,
                                const int32_t* __restrict__ r_done, int32_t stride,
                                const uint8_t* __restrict__ pool_adj,
                                const int64_t* __restrict__ pool_adj_off,
                                const int32_t* __restrict__ pool_n,
                                const int32_t* __restrict__ out_node_off,
                                const int64_t* __restrict__ out_adj_off, uint8_t* __restrict__ out_adj,
                                int32_t* __restrict__ out_cur, int32_t* __restrict__ out_next,
                                int32_t* __restrict__ out_action, float* __restrict__ out_reward,
                                int32_t* __restrict__ out_done, int32_t* __restrict__ err) {
    const int b = blockIdx.y;
    const int32_t s = slots[b];
    const int32_t gid = r_graph[s];
    const int32_t n = pool_n[gid];
    const int32_t v0 = out_node_off[b];
    if (out_node_off[b + 1] - v0 != n) {           // host offsets disagree with the ring: report, copy nothing
        if (threadIdx.x == 0 && blockIdx.x == 0 && err) atomicExch(err, b + 1);
        return;
    }
    const int64_t nn = (int64_t)n * n;
    const int64_t lo = (int64_t)blockIdx.x * GATHER_CHUNK;
    if (blockIdx.x == 0) {
        const int16_t* c = r_cur + (int64_t)s * stride;
        const int16_t* d = r_next + (int64_t)s * stride;
        for (int32_t i = threadIdx.x; i < n; i += blockDim.x) {
            out_cur[v0 + i] = c[i];
            out_next[v0 + i] = d[i];
        }
        if (threadIdx.x == 0) {
            out_action[b] = r_action[s];
            out_reward[b] = r_reward[s];
            out_done[b] = r_done[s];
        }
    }
    if (lo >= nn) return;
    const int64_t hi = min(nn, lo + GATHER_CHUNK);
    const uint8_t* src = pool_adj + pool_adj_off[gid];
    uint8_t* dst = out_adj + out_adj_off[b];
    if ((((uintptr_t)src ^ (uintptr_t)dst) & 15) == 0) {
        // same phase: bytes up to the first 16-byte boundary, uint4 body, byte tail
        int64_t a = lo + ((16 - ((uintptr_t)(src + lo) & 15)) & 15);
        if (a > hi) a = hi;
        for (int64_t i = lo + threadIdx.x; i < a; i += blockDim.x) dst[i] = src[i];
        const int64_t nv = (hi - a) >> 4;
        const uint4* s4 = reinterpret_cast<const uint4*>(src + a);
        uint4* d4 = reinterpret_cast<uint4*>(dst + a);
        for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) d4[i] = __ldg(s4 + i);
        for (int64_t i = a + (nv << 4) + threadIdx.x; i < hi; i += blockDim.x) dst[i] = src[i];
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = src[i];
    }
}

}  // namespace gcrl

using namespace gcrl;

extern "C" int gcrl_replay_store(int32_t n_envs, const int32_t* slot, const int32_t* graph_id,
                                 const int32_t* node_off, const int32_t* cur_colour,
                                 const int32_t* next_colour, const int32_t* action, const float* reward,
                                 const int32_t* done, int32_t stride, int32_t* r_graph, int16_t* r_cur,
                                 int16_t* r_next, int32_t* r_action, float* r_reward, int32_t* r_done,
                                 void* stream) {
    GCRL_CHECK_ARG(n_envs >= 0 && stride > 0);
    if (n_envs == 0) return GCRL_OK;
    GCRL_CHECK_ARG(slot && graph_id && node_off && cur_colour && next_colour && action && reward && done);
    GCRL_CHECK_ARG(r_graph && r_cur && r_next && r_action && r_reward && r_done);
    const int threads = 256;
    const int64_t blocks = ((int64_t)n_envs * 32 + threads - 1) / threads;
    k_replay_store<<<(unsigned)(blocks < 65535 ? blocks : 65535), threads, 0, (cudaStream_t)stream>>>(
        n_envs, slot, graph_id, node_off, cur_colour, next_colour, action, reward, done, stride, r_graph,
        r_cur, r_next, r_action, r_reward, r_done);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

extern "C" int gcrl_replay_gather(int32_t batch, const int32_t* slots, const int32_t* r_graph,
                                  const int16_t* r_cur, const int16_t* r_next, const int32_t* r_action,
                                  const float* r_reward, const int32_t* r_done, int32_t stride,
                                  const uint8_t* pool_adj, const int64_t* pool_adj_off,
                                  const int32_t* pool_n, int32_t n_max, const int32_t* out_node_off,
                                  const int64_t* out_adj_off, uint8_t* out_adj, int32_t* out_cur,
                                  int32_t* out_next, int32_t* out_action, float* out_reward,
                                  int32_t* out_done, int32_t* err, void* stream) {
    GCRL_CHECK_ARG(batch >= 0 && stride > 0 && n_max > 0 && n_max <= stride && batch <= 65535);
    if (batch == 0) return GCRL_OK;
    GCRL_CHECK_ARG(slots && r_graph && r_cur && r_next && r_action && r_reward && r_done);
    GCRL_CHECK_ARG(pool_adj && pool_adj_off && pool_n && out_node_off && out_adj_off);
    GCRL_CHECK_ARG(out_adj && out_cur && out_next && out_action && out_reward && out_done);
    const int64_t nn = (int64_t)n_max * n_max;
    dim3 grid((unsigned)((nn + GATHER_CHUNK - 1) / GATHER_CHUNK), (unsigned)batch);
    k_replay_gather<<<grid, 256, 0, (cudaStream_t)stream>>>(
        slots, r_graph, r_cur, r_next, r_action, r_reward, r_done, stride, pool_adj, pool_adj_off, pool_n,
        out_node_off, out_adj_off, out_adj, out_cur, out_next, out_action, out_reward, out_done, err);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}
