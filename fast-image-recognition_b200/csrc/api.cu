// api.cu — the C-ABI of libfir_b200.so (include/fir_b200.h): handles, staging, kernel orchestration.
// There is no CPU fallback in this file or anywhere below it: every compute entry point needs a CUDA
// device and fails with FIR_ERR_CUDA otherwise.
#include "fir_common.cuh"
#include "handles.hpp"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>

namespace fir {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) { g_last_error = msg; return code; }

int Workspace::reserve(size_t bytes) {
    off = 0;
    ++generation;                        // whatever the previous call left in the workspace is no longer addressable
    if (bytes <= cap) return FIR_OK;
    if (base) {
        FIR_CUDA_TRY(cudaStreamSynchronize(stream));
        FIR_CUDA_TRY(cudaFree(base));
        base = nullptr; cap = 0;
    }
    size_t want = bytes + bytes / 8 + (1 << 20);
    FIR_CUDA_TRY(cudaMalloc(&base, want));
    cap = want;
    return FIR_OK;
}
void* Workspace::take(size_t bytes) {
    size_t a = (off + 255) & ~(size_t)255;
    if (a + bytes > cap) return nullptr;
    off = a + bytes;
    return base + a;
}
void Workspace::release() {
    if (base) cudaFree(base);
    base = nullptr; cap = 0; off = 0;
}

static size_t al(size_t b) { return (b + 255) & ~(size_t)255; }

// Make `queries` (nq x d, host or device) available as zero-padded device rows [nq][dp].
static int stage_queries(fir_gallery* g, const float* queries, int64_t nq, int memspace, const float** dq) {
    const int d = g->d, dp = g->dp;
    if (memspace == FIR_DEVICE && d == dp) { *dq = queries; return FIR_OK; }
    float* buf = (float*)g->ws.take(sizeof(float) * (size_t)nq * dp);
    if (!buf) return fail(FIR_ERR_INTERNAL, "workspace underestimated (queries)");
    if (memspace == FIR_HOST) {
        if (d != dp) FIR_CUDA_TRY(cudaMemsetAsync(buf, 0, sizeof(float) * (size_t)nq * dp, g->stream));
        FIR_CUDA_TRY(cudaMemcpy2DAsync(buf, sizeof(float) * dp, queries, sizeof(float) * d, sizeof(float) * d, (size_t)nq,
                                       cudaMemcpyHostToDevice, g->stream));
    } else {
        FIR_TRY(launch_pad_rows(queries, nq, d, buf, dp, g->stream));
        g->stats.gpu_launches++;
    }
    *dq = buf;
    return FIR_OK;
}

// gallery splits of the exact tile kernel: enough blocks for four per SM.  The top-k mode merges one partial list per split
// (cap 64); the class modes fold through atomics, so a small batch may be cut much finer (256 queries x 1M rows: 592 blocks
// instead of 256 — the kernel is latency-bound below ~4 blocks per SM)
static int pick_nsplit(int64_t nq, int64_t n, int n_sm, int64_t cap = 64) {
    int64_t qblocks = ceil_div(nq, kExactTile);
    int64_t ntiles = ceil_div(n, kExactTile);
    int64_t want = ceil_div((int64_t)n_sm * 4, qblocks);
    want = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, ntiles), cap));
    return (int)want;
}

// exact CUDA-core top-k over (optionally) a subset of queries; results to device out_* [nq][k]
int exact_topk_device(fir_gallery* g, const float* dq, int64_t nq, int k, int d_end, const int32_t* qmap,
                      const int32_t* n_active, float* part_d, int32_t* part_i, int nsplit, float* od, int32_t* oi,
                      int64_t active_offset, int64_t active_cap) {
    ExactParams p{};
    p.q = dq; p.nq = nq; p.ldq = g->dp;
    p.x = g->rows; p.n = g->n; p.ldx = g->dp;
    p.labels = g->labels;
    p.d_end = d_end; p.k = k; p.mode = MODE_TOPK;
    p.nsplit = nsplit;
    p.tiles_per_split = ceil_div(ceil_div(g->n, kExactTile), nsplit);
    p.part_dist = part_d; p.part_idx = part_i;
    p.qmap = qmap; p.n_active = n_active; p.active_offset = active_offset; p.active_cap = active_cap;
    if (active_cap > 0) p.nq = std::min<int64_t>(nq, active_cap);       // grid covers one window of the active list
    { auto* ev = g->prof_begin(FIR_KERNEL_EXACT_TILES); int st_ = launch_exact_tiles(g->metric, p, g->stream); g->prof_end(ev); FIR_TRY(st_); }
    FIR_TRY(launch_merge_parts(part_d, part_i, nsplit, k, (int64_t)nsplit * k, active_cap > 0 ? std::min<int64_t>(nq, active_cap) : nq, k, g->index_offset, qmap, n_active, od, oi, g->stream,
                               active_offset, active_cap));
    g->stats.gpu_launches += 2;
    return FIR_OK;
}

// small-batch path: nq <= 8 queries → 8-row zero-padded query block + one streaming pass → dist[nq][n] in the workspace
static int stream_distances(fir_gallery* g, const float* queries, int64_t nq, int memspace, int d_end, float** dist_out) {
    const int dp = g->dp;
    float* qbuf = (float*)g->ws.take(sizeof(float) * (size_t)kStreamMaxQueries * dp);
    float* dist = (float*)g->ws.take(sizeof(float) * (size_t)nq * g->n);
    if (!qbuf || !dist) return fail(FIR_ERR_INTERNAL, "workspace underestimated (stream path)");
    FIR_CUDA_TRY(cudaMemsetAsync(qbuf, 0, sizeof(float) * (size_t)kStreamMaxQueries * dp, g->stream));
    FIR_CUDA_TRY(cudaMemcpy2DAsync(qbuf, sizeof(float) * dp, queries, sizeof(float) * g->d, sizeof(float) * g->d, (size_t)nq,
                                   memspace == FIR_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, g->stream));
    auto* ev = g->prof_begin(FIR_KERNEL_STREAM_DISTANCES);
    int st = launch_stream_distances(g->metric, (int)nq, qbuf, dp, g->rows, dp, g->n, d_end, dist, g->n, g->n_sm, g->stream);
    g->prof_end(ev);
    FIR_TRY(st);
    g->stats.gpu_launches++;
    *dist_out = dist;
    return FIR_OK;
}

}  // namespace fir

using namespace fir;

extern "C" {

const char* fir_last_error_string(void) { return g_last_error.c_str(); }
int fir_version(void) { return FIR_B200_VERSION; }

int fir_device_count(int* count) {
    if (!count) return fail(FIR_ERR_BAD_ARG, "count is null");
    *count = 0;
    FIR_CUDA_TRY(cudaGetDeviceCount(count));
    return FIR_OK;
}
int fir_set_device(int device) {
    FIR_CUDA_TRY(cudaSetDevice(device));
    return FIR_OK;
}

int fir_gallery_create(const float* rows, const int32_t* labels, int64_t n, int32_t d, int32_t metric, int32_t memspace,
                       int64_t index_offset, fir_gallery** out) {
    if (!out) return fail(FIR_ERR_BAD_ARG, "out is null");
    *out = nullptr;
    if (!rows || n <= 0 || d <= 0) return fail(FIR_ERR_BAD_ARG, "empty gallery (the reference returns -1 for every query; create no handle)");
    if (metric < FIR_L2 || metric > FIR_KL) return fail(FIR_ERR_BAD_ARG, "unknown metric");
    if (n + index_offset > (int64_t)std::numeric_limits<int32_t>::max()) return fail(FIR_ERR_UNSUPPORTED, "global gallery index exceeds int32");
    fir_gallery* g = new fir_gallery();
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { delete g; return fail(FIR_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e)); }
    g->device = dev;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) { delete g; return fail(FIR_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e)); }
    g->n_sm = prop.multiProcessorCount;
    g->cc_major = prop.major;
    g->n = n; g->d = d; g->dp = round_up(d, kRowPad); g->metric = metric; g->index_offset = index_offset;
    auto cleanup = [&](int code) { fir_gallery_destroy(g); return code; };
    e = cudaMalloc(&g->rows, sizeof(float) * (size_t)n * g->dp);
    if (e != cudaSuccess) return cleanup(fail(FIR_ERR_OOM, std::string("gallery rows: ") + cudaGetErrorString(e)));
    e = cudaMalloc(&g->labels, sizeof(int32_t) * (size_t)n);
    if (e != cudaSuccess) return cleanup(fail(FIR_ERR_OOM, std::string("gallery labels: ") + cudaGetErrorString(e)));
    cudaMemcpyKind kind = memspace == FIR_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (g->dp != d) {
        e = cudaMemset(g->rows, 0, sizeof(float) * (size_t)n * g->dp);
        if (e != cudaSuccess) return cleanup(fail(FIR_ERR_CUDA, cudaGetErrorString(e)));
    }
    e = cudaMemcpy2D(g->rows, sizeof(float) * g->dp, rows, sizeof(float) * d, sizeof(float) * d, (size_t)n, kind);
    if (e != cudaSuccess) return cleanup(fail(FIR_ERR_CUDA, std::string("gallery upload: ") + cudaGetErrorString(e)));
    // labels: class ids; n_classes = max+1 (host copy kept for DEM build / classmin sizing)
    g->h_labels.assign((size_t)n, 0);
    if (labels) {
        if (memspace == FIR_HOST) std::memcpy(g->h_labels.data(), labels, sizeof(int32_t) * (size_t)n);
        else {
            e = cudaMemcpy(g->h_labels.data(), labels, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) return cleanup(fail(FIR_ERR_CUDA, cudaGetErrorString(e)));
        }
    }
    int32_t mx = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (g->h_labels[i] < 0) return cleanup(fail(FIR_ERR_BAD_ARG, "negative class label"));
        mx = std::max(mx, g->h_labels[i]);
    }
    g->n_classes = mx + 1;
    e = cudaMemcpy(g->labels, g->h_labels.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cleanup(fail(FIR_ERR_CUDA, cudaGetErrorString(e)));
    *out = g;
    return FIR_OK;
}

int fir_gallery_destroy(fir_gallery* g) {
    if (!g) return FIR_OK;
    cudaStreamSynchronize(g->stream);
    if (g->rows) cudaFree(g->rows);
    if (g->labels) cudaFree(g->labels);
    if (g->tensor_buf) cudaFree(g->tensor_buf);
    if (g->tensor_buf_nat) cudaFree(g->tensor_buf_nat);
    if (g->prefix_norm2) cudaFree(g->prefix_norm2);
    if (g->d_cls_begin) cudaFree(g->d_cls_begin);
    if (g->d_stats) cudaFree(g->d_stats);
    if (g->d_l1max) cudaFree(g->d_l1max);
    if (g->kl_ent) cudaFree(g->kl_ent);
    for (auto& e : g->ev_pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    g->ws.release();
    delete g;
    return FIR_OK;
}

int fir_gallery_set_stream(fir_gallery* g, void* cuda_stream) {
    if (!g) return fail(FIR_ERR_BAD_ARG, "gallery is null");
    g->stream = (cudaStream_t)cuda_stream;
    g->ws.stream = g->stream;
    return FIR_OK;
}

int fir_gallery_info(const fir_gallery* g, int64_t* n, int32_t* d, int32_t* metric, int32_t* n_classes) {
    if (!g) return fail(FIR_ERR_BAD_ARG, "gallery is null");
    if (n) *n = g->n;
    if (d) *d = g->d;
    if (metric) *metric = g->metric;
    if (n_classes) *n_classes = g->n_classes;
    return FIR_OK;
}

int fir_gallery_index_offset(const fir_gallery* g, int64_t* index_offset) {
    if (!g || !index_offset) return fail(FIR_ERR_BAD_ARG, "null argument");
    *index_offset = g->index_offset;
    return FIR_OK;
}

int fir_gallery_set_num_classes(fir_gallery* g, int32_t n_classes) {
    if (!g) return fail(FIR_ERR_BAD_ARG, "gallery is null");
    int32_t mx = 0;
    for (int32_t l : g->h_labels) mx = std::max(mx, l);
    if (n_classes <= mx) return fail(FIR_ERR_BAD_ARG, "n_classes must exceed the largest label of the handle");
    g->n_classes = n_classes;
    return FIR_OK;
}

int fir_normalize_rows(float* rows, int64_t n, int32_t d, int32_t metric, int32_t memspace, void* cuda_stream) {
    if (!rows || n < 0 || d <= 0) return fail(FIR_ERR_BAD_ARG, "bad rows");
    if (metric < FIR_L2 || metric > FIR_NORM_VIDEO_SUMSQ) return fail(FIR_ERR_BAD_ARG, "unknown normalisation");
    if (n == 0) return FIR_OK;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    if (memspace == FIR_DEVICE) return launch_normalize_rows(rows, n, d, d, metric, s);
    float* buf = nullptr;
    FIR_CUDA_TRY(cudaMalloc(&buf, sizeof(float) * (size_t)n * d));
    int st = FIR_OK;
    cudaError_t e = cudaMemcpyAsync(buf, rows, sizeof(float) * (size_t)n * d, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) st = launch_normalize_rows(buf, n, d, d, metric, s);
    if (e == cudaSuccess && st == FIR_OK) e = cudaMemcpyAsync(rows, buf, sizeof(float) * (size_t)n * d, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(buf);
    if (st != FIR_OK) return st;
    if (e != cudaSuccess) return fail(FIR_ERR_CUDA, cudaGetErrorString(e));
    return FIR_OK;
}

int fir_search_topk(fir_gallery* g, const float* queries, int64_t nq, int32_t k, int32_t max_features, int32_t path,
                    int32_t memspace, int32_t* out_idx, float* out_dist) {
    if (!g) return fail(FIR_ERR_BAD_ARG, "gallery is null");
    if (nq < 0 || (nq > 0 && (!queries || !out_idx))) return fail(FIR_ERR_BAD_ARG, "bad query/output pointers");
    if (k < 1 || k > 1024) return fail(FIR_ERR_BAD_ARG, "k must be in [1,1024]");
    if (max_features < 0 || max_features > g->d) return fail(FIR_ERR_BAD_ARG, "max_features out of range");
    g->stats = fir_search_stats{};
    g->dbg_cand_idx = nullptr; g->dbg_cand_val = nullptr; g->dbg_cand_exact = nullptr;      // diagnostics of an earlier tensor call are stale now
    if (nq == 0) return FIR_OK;
    FIR_CUDA_TRY(cudaSetDevice(g->device));
    const int d_end = max_features > 0 ? max_features : g->d;
    bool tensor_ok = g->metric == FIR_L2 && tensor_path_supported(g->d) && g->cc_major >= 10 && k <= 28 && !(max_features > 0 && g->tensor_center);
    if (path == FIR_PATH_TENSOR && !tensor_ok)
        return fail(FIR_ERR_UNSUPPORTED, "tensor path needs L2, k<=28 and an sm_100 device");
    bool use_tensor = (path == FIR_PATH_TENSOR) || (path == FIR_PATH_AUTO && tensor_ok && nq * g->n >= (int64_t)1 << 22);
    if (use_tensor) return tensor_search_topk(g, queries, nq, k, memspace, out_idx, out_dist, d_end);
    const bool approx_ok = g->metric != FIR_L2 && max_features == 0 && k <= 28;
    if (path == FIR_PATH_APPROX && !approx_ok) return fail(FIR_ERR_UNSUPPORTED, "approximate path needs chi2/KL, all dimensions and k<=28");
    if (path == FIR_PATH_APPROX || (path == FIR_PATH_AUTO && approx_ok && nq > kStreamMaxQueries && nq * g->n >= (int64_t)1 << 22)) {
        const int st = approx_search_topk(g, queries, nq, k, memspace, out_idx, out_dist);
        if (st != kApproxDeclined) return st;                  // KL over a mixed-sign gallery: fall through to the exact kernels
    }

    if (nq <= kStreamMaxQueries && k <= kStreamMaxK && g->n >= 4096) {
        // latency mode: one pass over the gallery for all (<= 8) queries, then per-segment top-k + merge
        const int nseg = (int)std::max<int64_t>(1, std::min<int64_t>(1024, g->n / 2048));
        size_t need_s = al(sizeof(float) * (size_t)kStreamMaxQueries * g->dp) + al(sizeof(float) * (size_t)nq * g->n) +
                        2 * al((size_t)nq * nseg * k * 4) + 2 * al((size_t)nq * k * 4) + 4096;
        FIR_TRY(g->ws.reserve(need_s));
        float* dist = nullptr;
        FIR_TRY(stream_distances(g, queries, nq, memspace, d_end, &dist));
        float* part_d = (float*)g->ws.take((size_t)nq * nseg * k * 4);
        int32_t* part_i = (int32_t*)g->ws.take((size_t)nq * nseg * k * 4);
        float* od = out_dist; int32_t* oi = out_idx;
        if (memspace == FIR_HOST || !out_dist) od = (float*)g->ws.take((size_t)nq * k * 4);
        if (memspace == FIR_HOST) oi = (int32_t*)g->ws.take((size_t)nq * k * 4);
        if (!part_d || !part_i || !od || !oi) return fail(FIR_ERR_INTERNAL, "workspace underestimated (stream topk)");
        FIR_TRY(launch_stream_topk(dist, g->n, g->n, (int)nq, k, nseg, part_d, part_i, g->stream));
        FIR_TRY(launch_merge_parts(part_d, part_i, nseg, k, (int64_t)nseg * k, nq, k, g->index_offset, nullptr, nullptr, od, oi, g->stream));
        g->stats.gpu_launches += 2;
        g->stats.path_used = FIR_PATH_EXACT;
        if (memspace == FIR_HOST) {
            FIR_CUDA_TRY(cudaMemcpyAsync(out_idx, oi, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, g->stream));
            if (out_dist) FIR_CUDA_TRY(cudaMemcpyAsync(out_dist, od, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, g->stream));
            FIR_CUDA_TRY(cudaStreamSynchronize(g->stream));
        }
        return FIR_OK;
    }
    const int nsplit = pick_nsplit(nq, g->n, g->n_sm);
    size_t need = al(sizeof(float) * (size_t)nq * g->dp) + 2 * al((size_t)nq * nsplit * k * 4) + 2 * al((size_t)nq * k * 4) + 4096;
    FIR_TRY(g->ws.reserve(need));
    const float* dq = nullptr;
    FIR_TRY(stage_queries(g, queries, nq, memspace, &dq));
    float* part_d = (float*)g->ws.take((size_t)nq * nsplit * k * 4);
    int32_t* part_i = (int32_t*)g->ws.take((size_t)nq * nsplit * k * 4);
    float* od = out_dist; int32_t* oi = out_idx;
    if (memspace == FIR_HOST || !out_dist) od = (float*)g->ws.take((size_t)nq * k * 4);
    if (memspace == FIR_HOST) oi = (int32_t*)g->ws.take((size_t)nq * k * 4);
    if (!part_d || !part_i || !od || !oi) return fail(FIR_ERR_INTERNAL, "workspace underestimated (topk)");
    FIR_TRY(exact_topk_device(g, dq, nq, k, d_end, nullptr, nullptr, part_d, part_i, nsplit, od, oi));
    g->stats.path_used = FIR_PATH_EXACT;
    if (memspace == FIR_HOST) {
        FIR_CUDA_TRY(cudaMemcpyAsync(out_idx, oi, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, g->stream));
        if (out_dist) FIR_CUDA_TRY(cudaMemcpyAsync(out_dist, od, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, g->stream));
        FIR_CUDA_TRY(cudaStreamSynchronize(g->stream));
    }
    return FIR_OK;
}

int fir_profile_enable(fir_gallery* g, int32_t on) {
    if (!g) return fail(FIR_ERR_BAD_ARG, "gallery is null");
    g->profiling = on != 0;
    g->ev_used = 0;
    return FIR_OK;
}

int fir_profile_read(fir_gallery* g, int32_t kernel, double* total_ms, int32_t* launches) {
    if (!g) return fail(FIR_ERR_BAD_ARG, "gallery is null");
    FIR_CUDA_TRY(cudaStreamSynchronize(g->stream));
    double tot = 0; int cnt = 0;
    for (size_t i = 0; i < g->ev_used; ++i) {
        if (g->ev_pool[i].kind != kernel) continue;
        float ms = 0.f;
        FIR_CUDA_TRY(cudaEventElapsedTime(&ms, g->ev_pool[i].a, g->ev_pool[i].b));
        tot += ms; ++cnt;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = cnt;
    return FIR_OK;
}

int fir_search_last_stats(const fir_gallery* g, fir_search_stats* stats) {
    if (!g || !stats) return fail(FIR_ERR_BAD_ARG, "null argument");
    if (g->stats.path_used == FIR_PATH_APPROX && g->stats.n_fallback < 0 && g->d_l1max) {
        fir_gallery* m = const_cast<fir_gallery*>(g);
        int32_t nf = 0; float mb = 0.f;
        FIR_CUDA_TRY(cudaStreamSynchronize(m->stream));
        FIR_CUDA_TRY(cudaMemcpy(&nf, m->d_l1max + 4, 4, cudaMemcpyDeviceToHost));
        FIR_CUDA_TRY(cudaMemcpy(&mb, m->d_l1max + 5, 4, cudaMemcpyDeviceToHost));
        m->stats.n_fallback = nf; m->stats.reserved = (float)nf; m->stats.approx_err_bound = mb;
    }
    if (g->stats.path_used == FIR_PATH_TENSOR && g->stats.n_fallback < 0 && g->d_stats) {
        fir_gallery* m = const_cast<fir_gallery*>(g);      // device-side counters are fetched on demand
        int32_t nf = 0; float mb = 0.f;
        FIR_CUDA_TRY(cudaStreamSynchronize(m->stream));
        FIR_CUDA_TRY(cudaMemcpy(&nf, m->d_stats + 4, 4, cudaMemcpyDeviceToHost));     // queries that needed more than pass 1
        int32_t nfinal = 0;
        FIR_CUDA_TRY(cudaMemcpy(&nfinal, m->d_stats + 7, 4, cudaMemcpyDeviceToHost)); // of those, re-run exactly on CUDA cores
        m->stats.reserved = (float)nfinal;
        FIR_CUDA_TRY(cudaMemcpy(&mb, m->d_stats + 5, 4, cudaMemcpyDeviceToHost));
        m->stats.n_fallback = nf;
        m->stats.approx_err_bound = mb;
    }
    *stats = g->stats;
    return FIR_OK;
}

int fir_pair_distances(fir_gallery* g, const float* queries, int64_t nq, const int32_t* cand_idx, int32_t r, int32_t gallery_is_lhs,
                       int32_t memspace, float* out_dist) {
    if (!g) return fail(FIR_ERR_BAD_ARG, "gallery is null");
    if (nq < 0 || r < 1 || (nq > 0 && (!queries || !cand_idx || !out_dist))) return fail(FIR_ERR_BAD_ARG, "bad arguments");
    if (nq == 0) return FIR_OK;
    FIR_CUDA_TRY(cudaSetDevice(g->device));
    size_t need = al(sizeof(float) * (size_t)nq * g->dp) + 2 * al((size_t)nq * r * 4) + 4096;
    FIR_TRY(g->ws.reserve(need));
    const float* dq = nullptr;
    FIR_TRY(stage_queries(g, queries, nq, memspace, &dq));
    const int32_t* dc = cand_idx; float* od = out_dist;
    if (memspace == FIR_HOST) {
        int32_t* c = (int32_t*)g->ws.take((size_t)nq * r * 4);
        od = (float*)g->ws.take((size_t)nq * r * 4);
        if (!c || !od) return fail(FIR_ERR_INTERNAL, "workspace underestimated (pairs)");
        FIR_CUDA_TRY(cudaMemcpyAsync(c, cand_idx, (size_t)nq * r * 4, cudaMemcpyHostToDevice, g->stream));
        dc = c;
    }
    FIR_TRY(launch_pair_distances(g->metric, dq, nq, g->dp, g->rows, g->dp, g->n, g->d, dc, r, gallery_is_lhs, od, g->stream));
    if (memspace == FIR_HOST) {
        FIR_CUDA_TRY(cudaMemcpyAsync(out_dist, od, (size_t)nq * r * 4, cudaMemcpyDeviceToHost, g->stream));
        FIR_CUDA_TRY(cudaStreamSynchronize(g->stream));
    }
    return FIR_OK;
}

static int class_reduce(fir_gallery* g, const float* queries, int64_t nq, int memspace, int mode, double var, int64_t n_total,
                        float* out_min, int32_t* out_arg, double* out_scores, int32_t* out_label) {
    if (!g) return fail(FIR_ERR_BAD_ARG, "gallery is null");
    if (nq < 0 || (nq > 0 && !queries)) return fail(FIR_ERR_BAD_ARG, "bad query pointer");
    if (nq == 0) return FIR_OK;
    FIR_CUDA_TRY(cudaSetDevice(g->device));
    const int C = g->n_classes;
    const size_t cells = (size_t)nq * C;
    const bool stream = nq <= kStreamMaxQueries && g->n >= 4096;
    // Euclidean per-class nearest neighbour of a large batch: tcgen05 candidate passes + exact rerank (l2_tensor.cu)
    static const int cls_tensor_on = [] { const char* e = getenv("FIR_CLASSMIN_TENSOR"); return e ? atoi(e) : 1; }();
    if (mode == MODE_CLASSMIN && cls_tensor_on && g->metric == FIR_L2 && !stream && tensor_path_supported(g->d) && g->cc_major >= 10 && tensor_cta_mode() == 2 &&
        out_min && out_arg && nq * g->n >= (int64_t)1 << 24 && (int64_t)C * 4 <= 1 << 20) {
        const int st = tensor_class_min(g, queries, nq, memspace, out_min, out_arg);
        if (st != kApproxDeclined) return st;
    }
    size_t need = al(sizeof(float) * (size_t)nq * g->dp) + al(cells * 8) + al(cells * 4) * 2 + al((size_t)nq * 4) + 4096;
    if (stream) need += al(sizeof(float) * (size_t)kStreamMaxQueries * g->dp) + al(sizeof(float) * (size_t)nq * g->n);
    FIR_TRY(g->ws.reserve(need));
    const float* dq = nullptr;
    float* sdist = nullptr;
    if (stream) FIR_TRY(stream_distances(g, queries, nq, memspace, g->d, &sdist));
    else FIR_TRY(stage_queries(g, queries, nq, memspace, &dq));
    ExactParams p{};
    p.q = dq; p.nq = nq; p.ldq = g->dp;
    p.x = g->rows; p.n = g->n; p.ldx = g->dp;
    p.labels = g->labels; p.d_end = g->d; p.k = 0; p.mode = mode;
    p.nsplit = pick_nsplit(nq, g->n, g->n_sm, 1024);
    p.tiles_per_split = ceil_div(ceil_div(g->n, kExactTile), p.nsplit);
    p.n_classes = C;
    if (mode == MODE_CLASSMIN) {
        if (!out_min || !out_arg) return fail(FIR_ERR_BAD_ARG, "null outputs");
        unsigned long long* keys = (unsigned long long*)g->ws.take(cells * 8);
        float* dmin = out_min; int32_t* darg = out_arg;
        if (memspace == FIR_HOST) { dmin = (float*)g->ws.take(cells * 4); darg = (int32_t*)g->ws.take(cells * 4); }
        if (!keys || !dmin || !darg) return fail(FIR_ERR_INTERNAL, "workspace underestimated (classmin)");
        FIR_TRY(launch_fill_u64(keys, (int64_t)cells, ~0ull, g->stream));
        p.cls_key = keys;
        if (stream) FIR_TRY(launch_stream_class(sdist, g->n, g->n, (int)nq, g->labels, C, MODE_CLASSMIN, 0.0, keys, nullptr, g->stream));
        else { auto* ev = g->prof_begin(FIR_KERNEL_EXACT_TILES); const int st_ = launch_exact_tiles(g->metric, p, g->stream); g->prof_end(ev); FIR_TRY(st_); }
        FIR_TRY(launch_classmin_finalize(keys, (int64_t)cells, g->index_offset, dmin, darg, g->stream));
        if (memspace == FIR_HOST) {
            FIR_CUDA_TRY(cudaMemcpyAsync(out_min, dmin, cells * 4, cudaMemcpyDeviceToHost, g->stream));
            FIR_CUDA_TRY(cudaMemcpyAsync(out_arg, darg, cells * 4, cudaMemcpyDeviceToHost, g->stream));
            FIR_CUDA_TRY(cudaStreamSynchronize(g->stream));
        }
    } else {
        if (!(var > 0)) return fail(FIR_ERR_BAD_ARG, "var must be > 0");
        double* sc = out_scores; int32_t* lab = out_label;
        if (memspace == FIR_HOST || !out_scores) sc = (double*)g->ws.take(cells * 8);
        if (memspace == FIR_HOST && out_label) lab = (int32_t*)g->ws.take((size_t)nq * 4);
        if (!sc) return fail(FIR_ERR_INTERNAL, "workspace underestimated (pnn)");
        FIR_CUDA_TRY(cudaMemsetAsync(sc, 0, cells * 8, g->stream));
        p.cls_score = sc; p.two_var = 2 * var;
        if (stream) FIR_TRY(launch_stream_class(sdist, g->n, g->n, (int)nq, g->labels, C, MODE_PNN, 2 * var, nullptr, sc, g->stream));
        else { auto* ev = g->prof_begin(FIR_KERNEL_EXACT_TILES); const int st_ = launch_exact_tiles(g->metric, p, g->stream); g->prof_end(ev); FIR_TRY(st_); }
        FIR_TRY(launch_pnn_finalize(sc, nq, C, (double)(n_total > 0 ? n_total : g->n), lab, g->stream));
        if (memspace == FIR_HOST) {
            if (out_scores) FIR_CUDA_TRY(cudaMemcpyAsync(out_scores, sc, cells * 8, cudaMemcpyDeviceToHost, g->stream));
            if (out_label) FIR_CUDA_TRY(cudaMemcpyAsync(out_label, lab, (size_t)nq * 4, cudaMemcpyDeviceToHost, g->stream));
            FIR_CUDA_TRY(cudaStreamSynchronize(g->stream));
        }
    }
    return FIR_OK;
}

int fir_class_min(fir_gallery* g, const float* queries, int64_t nq, int32_t memspace, float* out_min, int32_t* out_arg) {
    return class_reduce(g, queries, nq, memspace, MODE_CLASSMIN, 0, 0, out_min, out_arg, nullptr, nullptr);
}

int fir_pnn_scores(fir_gallery* g, const float* queries, int64_t nq, double var, int64_t n_total, int32_t memspace,
                   double* out_scores, int32_t* out_label) {
    return class_reduce(g, queries, nq, memspace, MODE_PNN, var, n_total, nullptr, nullptr, out_scores, out_label);
}

static int twd_entry(fir_gallery* g, const float* queries, int64_t nq, int kind, int type, double threshold, int feat_count, int last_feature,
                     int memspace, int32_t* out_index, int32_t* out_label, uint8_t* out_unreliable) {
    if (!g) return fail(FIR_ERR_BAD_ARG, "gallery is null");
    if (nq < 0 || (nq > 0 && !queries)) return fail(FIR_ERR_BAD_ARG, "bad query pointer");
    if (memspace != FIR_HOST && memspace != FIR_DEVICE) return fail(FIR_ERR_BAD_ARG, "bad memspace");
    if (feat_count < 1 || last_feature < 1) return fail(FIR_ERR_BAD_ARG, "feat_count and last_feature must be positive");
    if (!(threshold == threshold)) return fail(FIR_ERR_BAD_ARG, "threshold is NaN");
    if (kind == 0) {
        if (type < 0 || type > 2) return fail(FIR_ERR_BAD_ARG, "type must be 0 (posteriors), 1 (diff) or 2 (ratio)");
        if (feat_count >= last_feature) return fail(FIR_ERR_BAD_ARG, "feat_count must be below last_feature");
        if (last_feature > g->d) return fail(FIR_ERR_BAD_ARG, "last_feature exceeds the gallery dimension");
        if (type == 0 && g->n_classes < 5) return fail(FIR_ERR_UNSUPPORTED, "the posterior test sums the five largest class posteriors: needs >= 5 classes");
    } else {
        if (threshold == 0) return fail(FIR_ERR_BAD_ARG, "threshold must be non-zero (the classifier uses 1/threshold)");
        // the last chunk may run past last_feature when feat_count does not divide it (ImageTesting.cpp:223,243)
        if ((int64_t)ceil_div(last_feature, feat_count) * feat_count > g->d) return fail(FIR_ERR_BAD_ARG, "the last chunk exceeds the gallery dimension");
    }
    if (nq == 0) return FIR_OK;
    FIR_CUDA_TRY(cudaSetDevice(g->device));
    int64_t mq = 0;
    const size_t need = al(sizeof(float) * (size_t)nq * g->dp) + twd_workspace_bytes(g, nq, &mq) + 2 * al(4 * (size_t)nq) + al((size_t)nq) + 4096;
    FIR_TRY(g->ws.reserve(need));
    const float* dq = nullptr;
    FIR_TRY(stage_queries(g, queries, nq, memspace, &dq));
    int32_t* di = out_index; int32_t* dl = out_label; unsigned char* du = out_unreliable;
    if (memspace == FIR_HOST) {
        di = out_index ? (int32_t*)g->ws.take(4 * (size_t)nq) : nullptr;
        dl = out_label ? (int32_t*)g->ws.take(4 * (size_t)nq) : nullptr;
        du = out_unreliable ? (unsigned char*)g->ws.take((size_t)nq) : nullptr;
        if ((out_index && !di) || (out_label && !dl) || (out_unreliable && !du)) return fail(FIR_ERR_INTERNAL, "workspace underestimated (twd outputs)");
    }
    FIR_TRY(twd_run(g, dq, nq, mq, kind, type, threshold, feat_count, last_feature, di, dl, du));
    if (memspace == FIR_HOST) {
        if (out_index) FIR_CUDA_TRY(cudaMemcpyAsync(out_index, di, 4 * (size_t)nq, cudaMemcpyDeviceToHost, g->stream));
        if (out_label) FIR_CUDA_TRY(cudaMemcpyAsync(out_label, dl, 4 * (size_t)nq, cudaMemcpyDeviceToHost, g->stream));
        if (out_unreliable) FIR_CUDA_TRY(cudaMemcpyAsync(out_unreliable, du, (size_t)nq, cudaMemcpyDeviceToHost, g->stream));
        FIR_CUDA_TRY(cudaStreamSynchronize(g->stream));
    }
    return FIR_OK;
}

int fir_twd_conventional(fir_gallery* g, const float* queries, int64_t nq, int32_t type, double threshold, int32_t feat_count,
                         int32_t last_feature, int32_t memspace, int32_t* out_index, int32_t* out_label, uint8_t* out_unreliable) {
    return twd_entry(g, queries, nq, 0, type, threshold, feat_count, last_feature, memspace, out_index, out_label, out_unreliable);
}

int fir_twd_proposed(fir_gallery* g, const float* queries, int64_t nq, int32_t feat_count, double threshold, int32_t last_feature,
                     int32_t memspace, int32_t* out_index, int32_t* out_label, uint8_t* out_unreliable) {
    return twd_entry(g, queries, nq, 1, 0, threshold, feat_count, last_feature, memspace, out_index, out_label, out_unreliable);
}

int fir_merge_topk(const float* parts_dist, const int32_t* parts_idx, int32_t n_parts, int64_t nq, int32_t k, float* out_dist,
                   int32_t* out_idx, void* cuda_stream) {
    if (!parts_dist || !parts_idx || !out_dist || !out_idx || n_parts < 1 || k < 1 || nq < 0) return fail(FIR_ERR_BAD_ARG, "bad arguments");
    return launch_merge_parts(parts_dist, parts_idx, n_parts, nq * k, k, nq, k, 0, nullptr, nullptr, out_dist, out_idx, (cudaStream_t)cuda_stream);
}

}  // extern "C"
