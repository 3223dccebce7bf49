// prof.cu -- instrumentation: kernel-launch counter and optional CUDA-event timing of the
// dominant kernels (per-kind totals), read by bench.py for the roofline figures.
#include "common.cuh"
#include <vector>
#include <mutex>

namespace gcrl {

static int64_t g_launches = 0;
static bool g_prof_on = false;
struct ProfRec { int kind; cudaEvent_t a, b; };
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static std::mutex g_mu;

void count_launch(int n) { g_launches += n; }

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GCRL_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

static cudaEvent_t take_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

int prof_begin(int kind, cudaStream_t st) {
    if (!g_prof_on) return -1;
    std::lock_guard<std::mutex> lk(g_mu);
    ProfRec r;
    r.kind = kind;
    r.a = take_event();
    r.b = take_event();
    cudaEventRecord(r.a, st);
    g_recs.push_back(r);
    return (int)g_recs.size() - 1;
}

void prof_end(int token, cudaStream_t st) {
    if (token < 0) return;
    std::lock_guard<std::mutex> lk(g_mu);
    cudaEventRecord(g_recs[token].b, st);
}

}  // namespace gcrl

using namespace gcrl;

extern "C" {

int64_t gcrl_launch_count(void) { return g_launches; }

int gcrl_prof_enable(int on) {
    g_prof_on = (on != 0);
    return GCRL_OK;
}

int gcrl_prof_read(double* ms, int64_t* launches, int n_kinds) {
    GCRL_CHECK_ARG(ms && launches && n_kinds >= 1);
    std::lock_guard<std::mutex> lk(g_mu);
    for (int i = 0; i < n_kinds; ++i) { ms[i] = 0.0; launches[i] = 0; }
    for (ProfRec& r : g_recs) {
        GCRL_CUDA(cudaEventSynchronize(r.b));
        float t = 0.f;
        GCRL_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
        if (r.kind >= 0 && r.kind < n_kinds) { ms[r.kind] += (double)t; launches[r.kind] += 1; }
        g_pool.push_back(r.a);
        g_pool.push_back(r.b);
    }
    g_recs.clear();
    return GCRL_OK;
}

}  // extern "C"
