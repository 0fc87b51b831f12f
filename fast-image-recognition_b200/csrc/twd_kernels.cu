// twd_kernels.cu — the sequential three-way-decision (TWD) classifiers of qt_cpp/ImageTesting.cpp on the GPU.
//
// Reference semantics:
//   ConventionalTWDClassifier::recognize (ImageTesting.cpp:108-186): distances over the first `feat_count` dimensions, a
//     reliability test on the best match (posterior share / distance difference / distance ratio against the *previous* best
//     of another class, :124-126), and for unreliable queries a refinement to `last_feature` = 256 dimensions (:166-181).
//   ProposedTWDClassifier::recognize (:207-288, CHECK_ALL_INSTANCES build): `feat_count`-dimension chunks accumulated in
//     double, instances farther than bestDist/th are dropped after every chunk, stop when only one class is left.
// Every chunk distance is feature_distance() over [lo,hi) (db_features.cpp:22-42): one thread owns a (query, row) pair and
// accumulates sequentially with separately rounded fp32 operations (dist_step), so the values the decisions are taken on
// carry the reference's bits.  A pair's running distance does not depend on what was pruned before, so every chunk is
// evaluated for the whole gallery for the queries still undecided (a compacted list) and pruning is recorded by parking the
// running distance at +inf; the per-query walks (argmin with strict '<' from 100000, the class-change runner-up chain,
// pruning and the variant count) are warp-per-query scans in gallery order.
#include "handles.hpp"
#include <algorithm>
#include <cstring>

namespace fir {

constexpr int TT = 64;      // tile: 64 queries x 64 gallery rows
constexpr int TK = 32;      // dims staged per slab
constexpr int TLD = TK + 4;  // 36 ⇒ LDS.128 conflict-free across 8 consecutive rows
enum { TWD_SET = 0, TWD_ACC = 1, TWD_MIX = 2 };

// D[q][j] (op)= feature_distance(query q, row j, lo, hi) for the queries on the active list
template <int METRIC>
__global__ void __launch_bounds__(256) twd_range_kernel(const float* __restrict__ q, int ldq, const int32_t* __restrict__ qlist,
                                                        const int32_t* __restrict__ n_active, const float* __restrict__ x, int ldx, int64_t n,
                                                        int lo, int hi, int mode, int feat_count, int last_feature, double* __restrict__ D) {
    const int na = *n_active;
    const int p0 = blockIdx.y * TT;
    if (p0 >= na) return;
    __shared__ __align__(16) float qs[TT * TLD];
    __shared__ __align__(16) float xs[TT * TLD];
    __shared__ int32_t qid[TT];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t x0 = (int64_t)blockIdx.x * TT;
    if (tid < TT) qid[tid] = p0 + tid < na ? qlist[p0 + tid] : -1;
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int k0 = lo; k0 < hi; k0 += TK) {
        for (int i = tid; i < TT * TK; i += 256) {
            const int r = i / TK, c = i - r * TK;
            const int64_t xi = x0 + r;
            const bool kin = k0 + c < hi;
            qs[r * TLD + c] = (qid[r] >= 0 && kin) ? q[(int64_t)qid[r] * ldq + k0 + c] : 0.f;
            xs[r * TLD + c] = (xi < n && kin) ? x[xi * ldx + k0 + c] : 0.f;
        }
        __syncthreads();
        const int kmax = min(TK, hi - k0);
        int kk = 0;
        for (; kk + 4 <= kmax; kk += 4) {
            float4 qa[4], xa[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) qa[a] = *reinterpret_cast<const float4*>(&qs[(ty + 16 * a) * TLD + kk]);
#pragma unroll
            for (int b = 0; b < 4; ++b) xa[b] = *reinterpret_cast<const float4*>(&xs[(tx + 16 * b) * TLD + kk]);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {                                               // lhs = query (ImageTesting.cpp:117,243)
                    dist_step<METRIC>(acc[a][b], qa[a].x, xa[b].x);
                    dist_step<METRIC>(acc[a][b], qa[a].y, xa[b].y);
                    dist_step<METRIC>(acc[a][b], qa[a].z, xa[b].z);
                    dist_step<METRIC>(acc[a][b], qa[a].w, xa[b].w);
                }
        }
        for (; kk < kmax; ++kk) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dist_step<METRIC>(acc[a][b], qs[(ty + 16 * a) * TLD + kk], xs[(tx + 16 * b) * TLD + kk]);
        }
        __syncthreads();
    }
    const float span = (float)(hi - lo);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int qi = qid[ty + 16 * a];
        if (qi < 0) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t xi = x0 + tx + 16 * b;
            if (xi >= n) continue;
            const float f = __fdiv_rn(acc[a][b], span);                                     // db_features.cpp:40
            double* cell = D + (int64_t)qi * n + xi;
            if (mode == TWD_SET) *cell = (double)f;                                         // :117
            else if (mode == TWD_ACC) *cell = __dadd_rn(*cell, (double)f);                  // :243  (+inf stays +inf: pruned)
            else {                                                                          // :174-175: double*int + float*int, / int
                const float rest = __fmul_rn(f, (float)(last_feature - feat_count));
                *cell = __ddiv_rn(__dadd_rn(__dmul_rn(*cell, (double)feat_count), (double)rest), (double)last_feature);
            }
        }
    }
}

__device__ __forceinline__ double pos_inf() { return __longlong_as_double(0x7ff0000000000000ll); }

// argmin with the reference's strict '<' from 100000 (lowest index on ties); warp-wide, result in every lane
__device__ __forceinline__ void warp_argmin_row(const double* __restrict__ row, int64_t n, int lane, double& bv, int& bi) {
    bv = 100000.0; bi = -1;
#pragma unroll 8
    for (int64_t j = lane; j < n; j += 32) {                   // unrolled: eight loads in flight per lane, compared in index order
        const double v = row[j];
        if (v < bv) { bv = v; bi = (int)j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (ov < bv || (ov == bv && (bi < 0 || oi < bi)))) { bv = ov; bi = oi; }
    }
}

// ProposedTWDClassifier, one chunk's decision (ImageTesting.cpp:224-283); warp per undecided query
__global__ void __launch_bounds__(128) twd_proposed_decide_kernel(double* __restrict__ D, int64_t n, const int32_t* __restrict__ labels,
                                                                  const int32_t* __restrict__ qlist, const int32_t* __restrict__ n_active,
                                                                  double threshold, int first_chunk, int32_t* __restrict__ best_idx,
                                                                  unsigned char* __restrict__ done, unsigned char* __restrict__ unreliable) {
    const int lane = threadIdx.x & 31;
    const int pos = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pos >= *n_active) return;
    const int q = qlist[pos];
    double* row = D + (int64_t)q * n;
    double bv; int bi;
    warp_argmin_row(row, n, lane, bv, bi);
    if (bi < 0) bi = best_idx[q];                                                            // nothing left below 100000: bestInd keeps its value
    const double thr = __dmul_rn(bv, threshold);                                             // :256
    const int best_class = bi >= 0 ? labels[bi] : -1;
    int others = 0;
    const double inf = pos_inf();
#pragma unroll 8
    for (int64_t j = lane; j < n; j += 32) {
        const double v = row[j];
        if (v == inf) continue;                                                              // instances_to_check[j] == 0
        if (v > thr) row[j] = inf;                                                           // :262-263
        else if (labels[j] != best_class) ++others;                                          // :264-265
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) others += __shfl_xor_sync(0xffffffffu, others, o);
    if (lane == 0) {
        best_idx[q] = bi;
        if (others == 0) done[q] = 1;                                                        // num_of_variants == 1, :279
        else if (first_chunk) unreliable[q] = 1;                                             // :281-282
    }
}

// rebuild the list of undecided queries (order is irrelevant: queries are independent)
__global__ void twd_compact_kernel(const unsigned char* __restrict__ done, int mq, int32_t* __restrict__ qlist, int32_t* __restrict__ n_active) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = q < mq && !done[q];
    const unsigned m = __ballot_sync(0xffffffffu, live);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(n_active, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) qlist[base + __popc(m & ((1u << lane) - 1))] = q;
}

__global__ void twd_iota_kernel(int32_t* __restrict__ qlist, int mq, int32_t* __restrict__ n_active) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < mq) qlist[q] = q;
    if (q == 0) *n_active = mq;
}

// ConventionalTWDClassifier, first pass decision (ImageTesting.cpp:110-165); warp per query.
// The running best is the lexicographic minimum (distance, index) of the prefix; an element updates it iff it is strictly
// below the prefix minimum, and secondBestDist takes the previous best's distance whenever that update changes class, so the
// final value comes from the last such update — found with a warp prefix-min scan in gallery order.
__global__ void __launch_bounds__(128) twd_conventional_decide_kernel(const double* __restrict__ D, int64_t n, const int32_t* __restrict__ labels,
                                                                      int mq, int type, double threshold, int n_classes,
                                                                      unsigned long long* __restrict__ probabs, int32_t* __restrict__ best_idx,
                                                                      unsigned char* __restrict__ unreliable, int32_t* __restrict__ qlist,
                                                                      int32_t* __restrict__ n_active) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= mq) return;
    const double* row = D + (int64_t)q * n;
    unsigned long long* pb = probabs + (int64_t)q * n_classes;
    const double inf = pos_inf();
    double cv = 100000.0; int ci = -1;                                                       // bestDist / bestInd, :111
    double second = 100000.0;
    for (int64_t base = 0; base < n; base += 32) {
        const int64_t j = base + lane;
        const double v = j < n ? row[j] : inf;
        const int lab = j < n ? labels[j] : -1;
        if (type == 0 && j < n) {
            const double p = exp(__dmul_rn(-v, 100.0));                                      // :119, DIST_WEIGHT = 100
            atomicMax(pb + lab, (unsigned long long)__double_as_longlong(p));                // :120-121 (p >= 0: bit order = value order)
        }
        double pv = v; int pi = (int)j;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double ov = __shfl_up_sync(0xffffffffu, pv, o);
            const int oi = __shfl_up_sync(0xffffffffu, pi, o);
            if (lane >= o && (ov < pv || (ov == pv && oi < pi))) { pv = ov; pi = oi; }
        }
        double ev = __shfl_up_sync(0xffffffffu, pv, 1);
        int ei = __shfl_up_sync(0xffffffffu, pi, 1);
        if (lane == 0 || cv <= ev) { ev = cv; ei = ci; }                                     // the carried best has the lower index
        const bool update = v < ev;                                                          // :122
        const bool change = update && ei >= 0 && labels[ei] != lab;                          // :123
        const unsigned cm = __ballot_sync(0xffffffffu, change);
        if (cm) second = __shfl_sync(0xffffffffu, ev, 31 - __clz(cm));                       // :124, the last one wins
        const double lv = __shfl_sync(0xffffffffu, pv, 31);
        const int li = __shfl_sync(0xffffffffu, pi, 31);
        if (lv < cv) { cv = lv; ci = li; }
    }
    bool reliable;
    if (type == 0) {
        __syncwarp();
        __threadfence();
        // :143-150 — the five largest class posteriors; libstdc++'s nth_element leaves them in an unspecified order, here
        // they are summed largest first
        double top[5] = {-1.0, -1.0, -1.0, -1.0, -1.0};
        if (lane == 0) {
            for (int c = 0; c < n_classes; ++c) {
                double p = __longlong_as_double((long long)__ldcg(pb + c));
                if (p > top[4]) {
                    top[4] = p;
#pragma unroll
                    for (int t = 4; t > 0; --t)
                        if (top[t] > top[t - 1]) { const double s = top[t]; top[t] = top[t - 1]; top[t - 1] = s; }
                }
            }
        }
        double sum = 0.0;
#pragma unroll
        for (int t = 0; t < 5; ++t) sum += top[t];
        const double max_probab = exp(__dmul_rn(-cv, 100.0)) / sum;                         // :127-128,151
        reliable = max_probab > threshold;
    } else if (type == 1) {
        reliable = __dsub_rn(second, cv) > threshold;                                        // :160
    } else {
        reliable = __ddiv_rn(cv, second) < threshold;                                        // :163
    }
    if (lane == 0) {
        best_idx[q] = ci;
        if (!reliable) {                                                                     // :166-167
            unreliable[q] = 1;
            qlist[atomicAdd(n_active, 1)] = q;
        }
    }
}

// refined pass of the conventional classifier (:169-181): plain argmin over the mixed distances
__global__ void __launch_bounds__(128) twd_argmin_kernel(const double* __restrict__ D, int64_t n, const int32_t* __restrict__ qlist,
                                                         const int32_t* __restrict__ n_active, int32_t* __restrict__ best_idx) {
    const int lane = threadIdx.x & 31;
    const int pos = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pos >= *n_active) return;
    const int q = qlist[pos];
    double bv; int bi;
    warp_argmin_row(D + (int64_t)q * n, n, lane, bv, bi);
    if (lane == 0) best_idx[q] = bi;
}

__global__ void twd_finalize_kernel(const int32_t* __restrict__ best_idx, const int32_t* __restrict__ labels, int mq, int64_t index_offset,
                                    int32_t* __restrict__ out_index, int32_t* __restrict__ out_label) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= mq) return;
    const int b = best_idx[q];
    if (out_index) out_index[q] = b >= 0 ? (int32_t)(b + index_offset) : -1;
    if (out_label) out_label[q] = b >= 0 ? labels[b] : -1;                                   // :183-185 / :285-287
}

template <int METRIC>
static void launch_range(fir_gallery* g, const float* dq, const int32_t* qlist, const int32_t* n_active, int mq, int lo, int hi, int mode,
                         int fc, int last, double* D) {
    dim3 grid((unsigned)ceil_div(g->n, TT), (unsigned)ceil_div(mq, TT));
    twd_range_kernel<METRIC><<<grid, 256, 0, g->stream>>>(dq, g->dp, qlist, n_active, g->rows, g->dp, g->n, lo, hi, mode, fc, last, D);
}

static int range_pass(fir_gallery* g, const float* dq, const int32_t* qlist, const int32_t* n_active, int mq, int lo, int hi, int mode, int fc,
                      int last, double* D) {
    if (g->metric == FIR_L2) launch_range<FIR_L2>(g, dq, qlist, n_active, mq, lo, hi, mode, fc, last, D);
    else if (g->metric == FIR_CHI2) launch_range<FIR_CHI2>(g, dq, qlist, n_active, mq, lo, hi, mode, fc, last, D);
    else launch_range<FIR_KL>(g, dq, qlist, n_active, mq, lo, hi, mode, fc, last, D);
    g->stats.gpu_launches++;
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}

static size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

size_t twd_workspace_bytes(const fir_gallery* g, int64_t nq, int64_t* mq_out) {
    // the running-distance matrix is the big item: keep it near 8 GiB and walk the queries in chunks
    int64_t mq = std::max<int64_t>(TT, ((int64_t)8 << 30) / (8 * std::max<int64_t>(g->n, 1)) / TT * TT);
    mq = std::min<int64_t>(mq, (nq + TT - 1) / TT * TT);
    mq = std::min<int64_t>(mq, (int64_t)65535 * TT);                                         // grid.y of the range kernel
    *mq_out = mq;
    return al256(sizeof(double) * (size_t)mq * g->n) + al256(8 * (size_t)mq * g->n_classes) + 3 * al256(4 * (size_t)mq) + 2 * al256((size_t)mq) + 4096;
}

// kind 0 conventional (type/threshold/feat_count), 1 proposed (feat_count/th); dq: device queries [nq][dp];
// d_index/d_label/d_unrel: device outputs (any may be null)
int twd_run(fir_gallery* g, const float* dq, int64_t nq, int64_t mq, int kind, int type, double threshold, int feat_count, int last_feature,
            int32_t* d_index, int32_t* d_label, unsigned char* d_unrel) {
    cudaStream_t s = g->stream;
    const int64_t n = g->n;
    const int C = g->n_classes;
    double* D = (double*)g->ws.take(sizeof(double) * (size_t)mq * n);
    unsigned long long* probabs = (unsigned long long*)g->ws.take(8 * (size_t)mq * C);
    int32_t* best = (int32_t*)g->ws.take(4 * (size_t)mq);
    int32_t* qlist = (int32_t*)g->ws.take(4 * (size_t)mq);
    int32_t* n_active = (int32_t*)g->ws.take(256);
    unsigned char* done = (unsigned char*)g->ws.take((size_t)mq);
    unsigned char* unrel = (unsigned char*)g->ws.take((size_t)mq);
    if (!D || !probabs || !best || !qlist || !n_active || !done || !unrel) return fail(FIR_ERR_INTERNAL, "workspace underestimated (twd)");
    for (int64_t q0 = 0; q0 < nq; q0 += mq) {
        const int m = (int)std::min<int64_t>(mq, nq - q0);
        const float* cq = dq + q0 * g->dp;
        const unsigned wblocks = (unsigned)ceil_div(m, 4), tblocks = (unsigned)ceil_div(m, 256);
        FIR_CUDA_TRY(cudaMemsetAsync(unrel, 0, (size_t)m, s));
        FIR_CUDA_TRY(cudaMemsetAsync(best, 0xFF, 4 * (size_t)m, s));
        twd_iota_kernel<<<tblocks, 256, 0, s>>>(qlist, m, n_active);
        if (kind == 1) {
            FIR_CUDA_TRY(cudaMemsetAsync(D, 0, sizeof(double) * (size_t)m * n, s));
            FIR_CUDA_TRY(cudaMemsetAsync(done, 0, (size_t)m, s));
            const double inv = 1.0 / threshold;                                              // ProposedTWDClassifier ctor, :191
            for (int cur = 0; cur < last_feature; cur += feat_count) {                       // :223
                FIR_TRY(range_pass(g, cq, qlist, n_active, m, cur, cur + feat_count, TWD_ACC, feat_count, last_feature, D));
                twd_proposed_decide_kernel<<<wblocks, 128, 0, s>>>(D, n, g->labels, qlist, n_active, inv, cur == 0 ? 1 : 0, best, done, unrel);
                if (cur + feat_count < last_feature) {
                    FIR_CUDA_TRY(cudaMemsetAsync(n_active, 0, 4, s));
                    twd_compact_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, s>>>(done, m, qlist, n_active);
                    g->stats.gpu_launches++;
                }
                g->stats.gpu_launches++;
            }
        } else {
            FIR_TRY(range_pass(g, cq, qlist, n_active, m, 0, feat_count, TWD_SET, feat_count, last_feature, D));
            if (type == 0) FIR_CUDA_TRY(cudaMemsetAsync(probabs, 0, 8 * (size_t)m * C, s));
            FIR_CUDA_TRY(cudaMemsetAsync(n_active, 0, 4, s));
            twd_conventional_decide_kernel<<<wblocks, 128, 0, s>>>(D, n, g->labels, m, type, threshold, C, probabs, best, unrel, qlist, n_active);
            FIR_TRY(range_pass(g, cq, qlist, n_active, m, feat_count, last_feature, TWD_MIX, feat_count, last_feature, D));
            twd_argmin_kernel<<<wblocks, 128, 0, s>>>(D, n, qlist, n_active, best);
            g->stats.gpu_launches += 2;
        }
        twd_finalize_kernel<<<tblocks, 256, 0, s>>>(best, g->labels, m, g->index_offset, d_index ? d_index + q0 : nullptr, d_label ? d_label + q0 : nullptr);
        g->stats.gpu_launches += 2;
        if (d_unrel) FIR_CUDA_TRY(cudaMemcpyAsync(d_unrel + q0, unrel, (size_t)m, cudaMemcpyDeviceToDevice, s));
        FIR_CUDA_TRY(cudaGetLastError());
    }
    return FIR_OK;
}

}  // namespace fir
