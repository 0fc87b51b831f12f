// cls_kernels.cu — the fp64 kNN / PNN classifiers of qt_cpp/classification.cpp on the GPU.
//
// Reference semantics (classification.cpp:116-170, 188-226): features are double, both operands are
// mean-centred first (normalize(), :103-105: x - avg, q - avg — two separate roundings), the squared
// distance is accumulated sequentially over the dimensions in fp64, kNN divides by D and walks the
// neighbours in ascending distance until a class owns K votes, PNN sums exp(-dist / (2*D*var)) per class
// in training-set order.  One thread owns one (query, training row) pair and uses separately rounded
// __dsub_rn/__dmul_rn/__dadd_rn, so the distances are bit-identical to the reference's.
#include "fir_common.cuh"
#include <algorithm>
#include <cstring>

struct fir_classifier {
    int device = 0;
    cudaStream_t stream = 0;
    int64_t n = 0;
    int64_t n_total = 0;         // PNN denominator: n, or the full training-set size when the rows are a reduced set (:393)
    int d = 0, n_classes = 0;
    double* xc = nullptr;        // [n][d] training rows, already centred: fl(x - avg)
    double* avg = nullptr;       // [d]
    int32_t* labels = nullptr;   // [n] class of each training row (class-major order)
    int32_t* cls_begin = nullptr; // [C+1] row range of every class (rows are class-major)
    bool class_major = true;
    fir::Workspace ws;
    // optional timing of the distance kernels (CUDA events on the classifier's stream)
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t> > ev_pool; size_t ev_used = 0;
    std::pair<cudaEvent_t, cudaEvent_t>* prof_begin() {
        if (!profiling) return nullptr;
        if (ev_used == ev_pool.size()) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); ev_pool.push_back(std::make_pair(a, b)); }
        auto* e = &ev_pool[ev_used++]; cudaEventRecord(e->first, stream); return e;
    }
    void prof_end(std::pair<cudaEvent_t, cudaEvent_t>* e) { if (e) cudaEventRecord(e->second, stream); }
};

namespace fir {

__global__ void centre_rows_kernel(const double* __restrict__ src, const double* __restrict__ avg, int64_t n, int d, double* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * d) return;
    dst[i] = __dsub_rn(src[i], avg[i % d]);                                     // classification.cpp:104
}

constexpr int CT = 64;     // tile: 64 queries x 64 training rows
constexpr int CK = 16;     // dims per stage
constexpr int CLD = CK + 1;

__device__ __forceinline__ void cls_cp_async8(void* smem, const void* gmem, bool ok) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(ok ? 8 : 0));
}

// dist[q][t] = sum_fi fl(fl(xc[t][fi] - qc[q][fi])^2), sequential in fi   (classification.cpp:127-142, 201-212)
// 64 x 64 tile, 4 x 4 pairs per thread, 16-dimension slabs through a 2-stage cp.async ring (row stride 17 doubles: the 16
// rows a half-warp reads fall in 16 different bank pairs).  FUSE_PNN: instead of storing the tile, one thread per query folds
// its 64 distances in training order into exp(-dist/den) run-length sums per class (rows are class-major) and adds each run
// to scores[q][class] — the Q x N matrix never reaches HBM (Parzen sum of classification.cpp:213).
template <bool FUSE_PNN>
__global__ void __launch_bounds__(256) cls_dist_kernel(const double* __restrict__ qc, int64_t nq, const double* __restrict__ xc, int64_t n,
                                                       int d, double* __restrict__ out, int k_lo, int k_hi, int accumulate,
                                                       const int32_t* __restrict__ labels, int n_classes, double den, double* __restrict__ scores) {
    extern __shared__ __align__(16) unsigned char cls_smem[];
    double* qs = reinterpret_cast<double*>(cls_smem);          // [2][CT][CLD]
    double* xs = qs + 2 * CT * CLD;                             // [2][CT][CLD]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t q0 = (int64_t)blockIdx.y * CT, x0 = (int64_t)blockIdx.x * CT;
    // dimensions [k_lo, k_hi); with `accumulate` the running sums continue from `out` (sequential PNN: one 32-dim chunk per
    // call — the same additions in the same order as one uninterrupted sum)
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t qi = q0 + ty + 16 * a, xi = x0 + tx + 16 * b;
            acc[a][b] = (!FUSE_PNN && accumulate && qi < nq && xi < n) ? out[qi * n + xi] : 0.0;
        }
    auto load = [&](int k0, int st) {
        for (int i = tid; i < CT * CK; i += 256) {
            const int r = i / CK, c = i - r * CK;
            const int64_t qi = q0 + r, xi = x0 + r;
            const bool kok = k0 + c < k_hi;
            cls_cp_async8(&qs[(st * CT + r) * CLD + c], qc + (qi < nq ? qi : 0) * d + (kok ? k0 + c : 0), qi < nq && kok);
            cls_cp_async8(&xs[(st * CT + r) * CLD + c], xc + (xi < n ? xi : 0) * d + (kok ? k0 + c : 0), xi < n && kok);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    const int nst = (k_hi - k_lo + CK - 1) / CK;
    load(k_lo, 0);
    for (int it = 0; it < nst; ++it) {
        const int st = it & 1, k0 = k_lo + it * CK;
        if (it + 1 < nst) load(k0 + CK, st ^ 1);
        else asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 1;\n" ::);
        __syncthreads();
        const double* qb = qs + st * CT * CLD;
        const double* xb = xs + st * CT * CLD;
        const int kmax = min(CK, k_hi - k0);
        for (int kk = 0; kk < kmax; ++kk) {
            double qa[4], xa[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) qa[a] = qb[(ty + 16 * a) * CLD + kk];
#pragma unroll
            for (int b = 0; b < 4; ++b) xa[b] = xb[(tx + 16 * b) * CLD + kk];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    double diff = __dsub_rn(xa[b], qa[a]);                      // :137 / :209  diff = x' - q'
                    acc[a][b] = __dadd_rn(acc[a][b], __dmul_rn(diff, diff));    // :141 / :211
                }
        }
        __syncthreads();
    }
    if (!FUSE_PNN) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int64_t qi = q0 + ty + 16 * a, xi = x0 + tx + 16 * b;
                if (qi < nq && xi < n) out[qi * n + xi] = acc[a][b];
            }
        return;
    }
    // fused Parzen epilogue: the 64 x 64 distances go through shared memory (the operand ring is free now) to one thread per query
    double* ds = reinterpret_cast<double*>(cls_smem);           // [CT][CT + 1]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) ds[(ty + 16 * a) * (CT + 1) + tx + 16 * b] = acc[a][b];
    __syncthreads();
    if (tid < CT && q0 + tid < nq) {
        const int jmax = (int)min((int64_t)CT, n - x0);
        int cur_c = -1; double cur_sum = 0.0;
        for (int j = 0; j < jmax; ++j) {
            const int c = labels[x0 + j];
            const double e = exp(-ds[tid * (CT + 1) + j] / den);                // :213
            if (c != cur_c) {
                if (cur_c >= 0) atomicAdd(&scores[(q0 + tid) * n_classes + cur_c], cur_sum);
                cur_c = c; cur_sum = e;
            } else cur_sum += e;
        }
        if (cur_c >= 0) atomicAdd(&scores[(q0 + tid) * n_classes + cur_c], cur_sum);
    }
}
constexpr size_t kClsSmem = sizeof(double) * (size_t)(4 * CT * CLD) > sizeof(double) * (size_t)(CT * (CT + 1)) ? sizeof(double) * (size_t)(4 * CT * CLD)
                                                                                                           : sizeof(double) * (size_t)(CT * (CT + 1));

// scores /= n_total (:215)
__global__ void scale_scores_kernel(double* __restrict__ scores, int64_t cells, double n_total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cells) scores[i] = scores[i] / n_total;
}

// PNN: one thread per (query, class); sums in training-set order (classification.cpp:195-216)
__global__ void pnn_class_sum_kernel(const double* __restrict__ dist, int64_t nq, int64_t n, int n_classes, const int32_t* __restrict__ cls_begin,
                                     const int32_t* __restrict__ labels, int class_major, double den, double n_total, double* __restrict__ scores) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * n_classes) return;
    const int64_t q = i / n_classes;
    const int c = (int)(i - q * n_classes);
    const double* dr = dist + q * n;
    double s = 0.0;
    if (class_major) {
        for (int t = cls_begin[c]; t < cls_begin[c + 1]; ++t) s += exp(-dr[t] / den);      // :213
    } else {
        for (int64_t t = 0; t < n; ++t)
            if (labels[t] == c) s += exp(-dr[t] / den);
    }
    scores[i] = s / n_total;                                                               // :215
}

__global__ void argmax_double_kernel(const double* __restrict__ scores, int64_t nq, int n_classes, int32_t* __restrict__ label) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    double mx = -1.7976931348623157e308; int best = -1;                                    // :217-225
    for (int c = 0; c < n_classes; ++c) {
        double v = scores[q * n_classes + c];
        if (mx < v) { mx = v; best = c; }
    }
    label[q] = best;
}

// kNN vote: one warp per query.  Neighbours are produced in ascending (dist/D, index) order by repeated
// warp-wide "smallest pair greater than the last one" scans; lane 0 counts votes until a class reaches K
// (classification.cpp:143, 151-160), then takes the arg-max over classes with strict '<' (:161-169).
__global__ void __launch_bounds__(128) knn_vote_kernel(const double* __restrict__ dist, int64_t nq, int64_t n, int d, int n_classes,
                                                       const int32_t* __restrict__ labels, int K, float* __restrict__ votes_scratch,
                                                       int32_t* __restrict__ out_label) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const double* dr = dist + q * n;
    float* votes = votes_scratch + q * n_classes;
    for (int c = lane; c < n_classes; c += 32) votes[c] = 0.f;
    __syncwarp();
    const double dd = (double)(size_t)d;
    double last_d = -1.0; int64_t last_i = -1;
    bool done = false;
    for (int64_t step = 0; step < n && !done; ++step) {
        double bd = 0.0; int64_t bi = -1;
        for (int64_t t = lane; t < n; t += 32) {
            double v = dr[t] / dd;                                                         // :143
            if (step > 0 && !(v > last_d || (v == last_d && t > last_i))) continue;
            if (bi < 0 || v < bd || (v == bd && t < bi)) { bd = v; bi = t; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            double od = __shfl_xor_sync(0xffffffffu, bd, o);
            long long oi = __shfl_xor_sync(0xffffffffu, (long long)bi, o);
            if (oi >= 0 && (bi < 0 || od < bd || (od == bd && oi < bi))) { bd = od; bi = oi; }
        }
        if (bi < 0) break;
        last_d = bd; last_i = bi;
        if (lane == 0) {
            int c = labels[bi];
            float v = votes[c] + 1.f;                                                      // :156
            votes[c] = v;
            done = v >= (float)K;                                                          // :157
        }
        done = __shfl_sync(0xffffffffu, (int)done, 0) != 0;
    }
    __syncwarp();
    if (lane == 0) {
        float mx = -__int_as_float(0x7f800000); int best = -1;                             // (float)-DBL_MAX == -inf
        for (int c = 0; c < n_classes; ++c)
            if (mx < votes[c]) { mx = votes[c]; best = c; }
        out_label[q] = best;
    }
}

// predict_sequentional's per-chunk decision (classification.cpp:270-292), one thread per query: arg-max over the classes
// still checked, drop those below max/1e9 (threshold rounded to float as in the reference), stop when one class is left.
__global__ void pnn_seq_decide_kernel(const double* __restrict__ scores, int64_t nq, int n_classes, unsigned char* __restrict__ check,
                                      int32_t* __restrict__ best, unsigned char* __restrict__ done) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq || done[q]) return;
    const double* sc = scores + q * n_classes;
    unsigned char* ck = check + q * n_classes;
    double mx = -1.7976931348623157e308; int b = best[q];
    for (int c = 0; c < n_classes; ++c)
        if (ck[c] && mx < sc[c]) { mx = sc[c]; b = c; }
    best[q] = b;
    const float thr = (float)(mx / 1000000000.0);
    int variants = 0;
    for (int c = 0; c < n_classes; ++c)
        if (ck[c]) { if (sc[c] < (double)thr) ck[c] = 0; else ++variants; }
    if (variants == 1) done[q] = 1;
}

}  // namespace fir

using namespace fir;

extern "C" {

int fir_classifier_create(const double* train_rows, const int32_t* train_labels, int64_t n, int32_t d, int32_t n_classes,
                          const double* avg, fir_classifier** out) {
    if (!out) return fail(FIR_ERR_BAD_ARG, "out is null");
    *out = nullptr;
    if (!train_rows || !train_labels || !avg || n <= 0 || d <= 0 || n_classes <= 0) return fail(FIR_ERR_BAD_ARG, "bad training set");
    fir_classifier* c = new fir_classifier();
    cudaError_t e = cudaGetDevice(&c->device);
    if (e != cudaSuccess) { delete c; return fail(FIR_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e)); }
    c->n = n; c->d = d; c->n_classes = n_classes;
    std::vector<int32_t> begin((size_t)n_classes + 1, 0);
    bool major = true;
    for (int64_t t = 0; t < n; ++t) {
        int32_t l = train_labels[t];
        if (l < 0 || l >= n_classes) { delete c; return fail(FIR_ERR_BAD_ARG, "training label out of range"); }
        if (t > 0 && l < train_labels[t - 1]) major = false;
        begin[(size_t)l + 1]++;
    }
    for (int k = 0; k < n_classes; ++k) begin[(size_t)k + 1] += begin[(size_t)k];
    c->class_major = major;
    double* raw = nullptr;
    auto cleanup = [&](int code) { if (raw) cudaFree(raw); fir_classifier_destroy(c); return code; };
    if (cudaMalloc(&c->xc, sizeof(double) * (size_t)n * d) != cudaSuccess || cudaMalloc(&raw, sizeof(double) * (size_t)n * d) != cudaSuccess ||
        cudaMalloc(&c->avg, sizeof(double) * d) != cudaSuccess || cudaMalloc(&c->labels, sizeof(int32_t) * (size_t)n) != cudaSuccess ||
        cudaMalloc(&c->cls_begin, sizeof(int32_t) * ((size_t)n_classes + 1)) != cudaSuccess)
        return cleanup(fail(FIR_ERR_OOM, "classifier allocation failed"));
    if (cudaMemcpy(raw, train_rows, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->avg, avg, sizeof(double) * d, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->labels, train_labels, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->cls_begin, begin.data(), sizeof(int32_t) * ((size_t)n_classes + 1), cudaMemcpyHostToDevice) != cudaSuccess)
        return cleanup(fail(FIR_ERR_CUDA, "classifier upload failed"));
    centre_rows_kernel<<<(unsigned)ceil_div(n * d, 256), 256>>>(raw, c->avg, n, d, c->xc);
    e = cudaDeviceSynchronize();
    cudaFree(raw); raw = nullptr;
    if (e != cudaSuccess) return cleanup(fail(FIR_ERR_CUDA, std::string("centre_rows_kernel: ") + cudaGetErrorString(e)));
    *out = c;
    return FIR_OK;
}

int fir_classifier_set_total(fir_classifier* c, int64_t n_total) {
    if (!c || n_total < 0) return fail(FIR_ERR_BAD_ARG, "bad arguments");
    c->n_total = n_total;
    return FIR_OK;
}

int fir_classifier_destroy(fir_classifier* c) {
    if (!c) return FIR_OK;
    if (c->xc) cudaFree(c->xc);
    if (c->avg) cudaFree(c->avg);
    if (c->labels) cudaFree(c->labels);
    if (c->cls_begin) cudaFree(c->cls_begin);
    for (auto& e : c->ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    c->ws.release();
    delete c;
    return FIR_OK;
}

// shared driver: distances for a chunk of queries, then the requested reducer
static int classify(fir_classifier* c, const double* queries, int64_t nq, int K, bool pnn, double* out_scores, int32_t* out_label,
                    bool sequential = false, int memspace = FIR_HOST) {
    if (!c) return fail(FIR_ERR_BAD_ARG, "classifier is null");
    if (nq < 0 || (nq > 0 && (!queries || !out_label))) return fail(FIR_ERR_BAD_ARG, "bad arguments");
    if (!pnn && K < 1) return fail(FIR_ERR_BAD_ARG, "K must be >= 1");
    if (memspace != FIR_HOST && memspace != FIR_DEVICE) return fail(FIR_ERR_BAD_ARG, "bad memspace");
    if (nq == 0) return FIR_OK;
    FIR_CUDA_TRY(cudaSetDevice(c->device));
    const int64_t n = c->n; const int d = c->d, C = c->n_classes;
    const bool host = memspace == FIR_HOST;
    const bool fused = pnn && !sequential;                       // Parzen sums in the distance epilogue: no Q x N matrix (any row order: runs are per label)
    const int64_t chunk = fused ? std::max<int64_t>(64, std::min<int64_t>(nq, ((int64_t)256 << 20) / (8 * (int64_t)std::max(C, d))))
                                : std::max<int64_t>(64, std::min<int64_t>(nq, ((int64_t)512 << 20) / (8 * n)));   // <= 512 MiB of distances
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t need = 2 * al(sizeof(double) * (size_t)chunk * d) + (fused ? 0 : al(sizeof(double) * (size_t)chunk * n)) + al(sizeof(double) * (size_t)chunk * C) +
                  al(sizeof(float) * (size_t)chunk * C) + al(sizeof(int32_t) * (size_t)chunk) + al((size_t)chunk * C) + al((size_t)chunk) + 4096;
    FIR_TRY(c->ws.reserve(need));
    double* qraw = (double*)c->ws.take(sizeof(double) * (size_t)chunk * d);
    double* qc = (double*)c->ws.take(sizeof(double) * (size_t)chunk * d);
    double* dist = fused ? nullptr : (double*)c->ws.take(sizeof(double) * (size_t)chunk * n);
    double* sc = (double*)c->ws.take(sizeof(double) * (size_t)chunk * C);
    float* votes = (float*)c->ws.take(sizeof(float) * (size_t)chunk * C);
    int32_t* lab = (int32_t*)c->ws.take(sizeof(int32_t) * (size_t)chunk);
    unsigned char* check = (unsigned char*)c->ws.take((size_t)chunk * C);
    unsigned char* done = (unsigned char*)c->ws.take((size_t)chunk);
    if (!check || !done) return fail(FIR_ERR_INTERNAL, "workspace underestimated (classifier)");
    if (!qraw || !qc || (!fused && !dist) || !sc || !votes || !lab) return fail(FIR_ERR_INTERNAL, "workspace underestimated (classifier)");
    double var = 0.00002;                                       // classification.cpp:190
    if (d > 2000) var /= 10;                                    // :192-193
    const double den = (double)(size_t)(2 * (size_t)d) * var;   // :213  2*num_of_cont_features*var
    const double n_total = (double)(c->n_total > 0 ? c->n_total : n);
    cudaStream_t s = c->stream;
    for (int64_t lo = 0; lo < nq; lo += chunk) {
        const int64_t m = std::min(chunk, nq - lo);
        const double* qsrc = queries + lo * d;
        if (host) { FIR_CUDA_TRY(cudaMemcpyAsync(qraw, qsrc, sizeof(double) * (size_t)m * d, cudaMemcpyHostToDevice, s)); qsrc = qraw; }
        centre_rows_kernel<<<(unsigned)ceil_div(m * d, 256), 256, 0, s>>>(qsrc, c->avg, m, d, qc);
        dim3 grid((unsigned)ceil_div(n, CT), (unsigned)ceil_div(m, CT));
        if (fused) {
            FIR_CUDA_TRY(cudaMemsetAsync(sc, 0, sizeof(double) * (size_t)m * C, s));
            auto* ev = c->prof_begin();
            cls_dist_kernel<true><<<grid, 256, kClsSmem, s>>>(qc, m, c->xc, n, d, nullptr, 0, d, 0, c->labels, C, den, sc);
            c->prof_end(ev);
            scale_scores_kernel<<<(unsigned)ceil_div(m * C, 256), 256, 0, s>>>(sc, m * C, n_total);
            int32_t* lab_dst = host ? lab : out_label + lo;
            argmax_double_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, s>>>(sc, m, C, lab_dst);
            if (out_scores) FIR_CUDA_TRY(cudaMemcpyAsync(out_scores + lo * C, sc, sizeof(double) * (size_t)m * C, host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s));
        } else if (sequential) {
            // PNNClassifier::predict_sequentional: 32-dimension chunks (delta_features_count, classification.cpp:182).  A class that
            // is still checked gets the same score as in the reference whatever was pruned before, so every chunk is evaluated
            // for all classes and the pruning walk is replayed per query by pnn_seq_decide_kernel.
            FIR_CUDA_TRY(cudaMemsetAsync(check, 1, (size_t)m * C, s));
            FIR_CUDA_TRY(cudaMemsetAsync(done, 0, (size_t)m, s));
            FIR_CUDA_TRY(cudaMemsetAsync(lab, 0xFF, sizeof(int32_t) * (size_t)m, s));
            for (int cur = 0; cur < d; cur += 32) {
                const int max_fi = std::min(cur + 32, d);
                cls_dist_kernel<false><<<grid, 256, kClsSmem, s>>>(qc, m, c->xc, n, d, dist, cur, max_fi, cur > 0 ? 1 : 0, nullptr, 0, 0.0, nullptr);
                pnn_class_sum_kernel<<<(unsigned)ceil_div(m * C, 128), 128, 0, s>>>(dist, m, n, C, c->cls_begin, c->labels, c->class_major ? 1 : 0,
                                                                                   (2 * var) * (double)(size_t)max_fi, n_total, sc);
                pnn_seq_decide_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, s>>>(sc, m, C, check, lab, done);
            }
            if (!host) FIR_CUDA_TRY(cudaMemcpyAsync(out_label + lo, lab, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToDevice, s));
        } else {
            auto* ev = c->prof_begin();
            cls_dist_kernel<false><<<grid, 256, kClsSmem, s>>>(qc, m, c->xc, n, d, dist, 0, d, 0, nullptr, 0, 0.0, nullptr);
            c->prof_end(ev);
            if (pnn) {                                            // (not class-major: the per-class gather below walks the label array)
                pnn_class_sum_kernel<<<(unsigned)ceil_div(m * C, 128), 128, 0, s>>>(dist, m, n, C, c->cls_begin, c->labels, 0, den, n_total, sc);
                argmax_double_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, s>>>(sc, m, C, host ? lab : out_label + lo);
                if (out_scores) FIR_CUDA_TRY(cudaMemcpyAsync(out_scores + lo * C, sc, sizeof(double) * (size_t)m * C, host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s));
            } else knn_vote_kernel<<<(unsigned)ceil_div(m, 4), 128, 0, s>>>(dist, m, n, d, C, c->labels, K, votes, host ? lab : out_label + lo);
        }
        FIR_CUDA_TRY(cudaGetLastError());
        if (host) {
            FIR_CUDA_TRY(cudaMemcpyAsync(out_label + lo, lab, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost, s));
            FIR_CUDA_TRY(cudaStreamSynchronize(s));
        }
    }
    return FIR_OK;
}

int fir_classifier_knn(fir_classifier* c, const double* queries, int64_t nq, int32_t K, int32_t* out_label) {
    return classify(c, queries, nq, K, false, nullptr, out_label);
}

int fir_classifier_pnn_sequential(fir_classifier* c, const double* queries, int64_t nq, int32_t* out_label) {
    return classify(c, queries, nq, 0, true, nullptr, out_label, true);
}

int fir_classifier_pnn(fir_classifier* c, const double* queries, int64_t nq, double* out_scores, int32_t* out_label) {
    return classify(c, queries, nq, 0, true, out_scores, out_label);
}

// the same with an explicit memory space for queries / outputs (FIR_DEVICE: asynchronous on the classifier's stream)
int fir_classifier_knn_ex(fir_classifier* c, const double* queries, int64_t nq, int32_t K, int32_t memspace, int32_t* out_label) {
    return classify(c, queries, nq, K, false, nullptr, out_label, false, memspace);
}
int fir_classifier_pnn_ex(fir_classifier* c, const double* queries, int64_t nq, int32_t memspace, double* out_scores, int32_t* out_label) {
    return classify(c, queries, nq, 0, true, out_scores, out_label, false, memspace);
}
int fir_classifier_set_stream(fir_classifier* c, void* cuda_stream) {
    if (!c) return fail(FIR_ERR_BAD_ARG, "classifier is null");
    c->stream = (cudaStream_t)cuda_stream; c->ws.stream = c->stream;
    return FIR_OK;
}
// CUDA-event timing of the fp64 distance kernel (fused Parzen / kNN matrix) since profiling was switched on
int fir_classifier_profile(fir_classifier* c, int32_t on, double* total_ms, int32_t* launches) {
    if (!c) return fail(FIR_ERR_BAD_ARG, "classifier is null");
    if (total_ms || launches) {
        FIR_CUDA_TRY(cudaStreamSynchronize(c->stream));
        double tot = 0;
        for (size_t i = 0; i < c->ev_used; ++i) { float ms = 0.f; FIR_CUDA_TRY(cudaEventElapsedTime(&ms, c->ev_pool[i].first, c->ev_pool[i].second)); tot += ms; }
        if (total_ms) *total_ms = tot;
        if (launches) *launches = (int32_t)c->ev_used;
    }
    c->profiling = on != 0;
    c->ev_used = 0;
    return FIR_OK;
}

}  // extern "C"
