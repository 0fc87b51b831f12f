// kmedoids.cu — the training step of PNNwithClusteringClassifier (qt_cpp/classification.cpp:321-388) on the GPU.
//
// Reference semantics: every class with more than `num_clusters` training rows is reduced to `num_clusters` medoids by 100
// rounds of { assign each row to the nearest current medoid (strict '<' over clusters in order: the lowest cluster wins a
// tie) ; per cluster pick the member whose summed distance to the members is smallest (strict '<': the lowest row wins) },
// starting from the first `num_clusters` rows; distances are mean squared differences of the RAW rows accumulated
// sequentially in fp64.  Smaller classes are kept whole.  predict() (:389-428) is then the Parzen PNN over the kept rows
// with the FULL training-set size as denominator — fir_classifier_create over the selected rows + fir_classifier_set_total.
// A cluster that runs empty makes the reference index with (size_t)-1; here that is reported as an error.
#include "fir_common.cuh"
#include <algorithm>
#include <vector>

namespace fir {

struct KmedClass { int64_t row0; int64_t m_off; int32_t m; int32_t out_off; };     // first row, offset of its matrix, size, first output slot

// M_c[u][v] = (1/d) Σ_fi fl(fl(x_u − x_v)²), sequential in fi (:339-344, :361-366); one thread per pair, blockIdx.y = class
__global__ void kmed_matrix_kernel(const double* __restrict__ rows, int d, const KmedClass* __restrict__ cls, double* __restrict__ Mall) {
    const KmedClass kc = cls[blockIdx.y];
    const int m = kc.m;
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < (int64_t)m * m; id += (int64_t)gridDim.x * blockDim.x) {
        const int u = (int)(id / m), v = (int)(id % m);
        const double* a = rows + (kc.row0 + u) * d;
        const double* b = rows + (kc.row0 + v) * d;
        double dist = 0.0;
        for (int fi = 0; fi < d; ++fi) {
            const double diff = __dsub_rn(a[fi], b[fi]);
            dist = __dadd_rn(dist, __dmul_rn(diff, diff));
        }
        Mall[kc.m_off + id] = __ddiv_rn(dist, (double)d);
    }
}

// one block per class; best[]/sums[] live in global scratch (indexed by training row), the per-cluster arg-min goes through
// shared memory
__global__ void __launch_bounds__(256) kmed_iterate_kernel(const double* __restrict__ Mall, const KmedClass* __restrict__ cls, int K, int steps,
                                                           int32_t* __restrict__ best_all, double* __restrict__ sums_all,
                                                           int32_t* __restrict__ cent_all, int32_t* __restrict__ dead_all) {
    __shared__ double s_v[256];
    __shared__ int s_i[256];
    __shared__ int s_dead;
    const KmedClass kc = cls[blockIdx.x];
    const int m = kc.m;
    const double* M = Mall + kc.m_off;
    int32_t* best = best_all + kc.row0;
    double* sums = sums_all + kc.row0;
    int32_t* cent = cent_all + kc.out_off;
    const int tid = threadIdx.x;
    for (int c = tid; c < K; c += blockDim.x) cent[c] = c;                                    // :329-330
    if (tid == 0) s_dead = 0;
    __syncthreads();
    for (int step = 0; step < steps; ++step) {
        for (int t = tid; t < m; t += blockDim.x) {                                           // :333-352
            int b = -1; double bd = 1.7976931348623157e308;
            for (int c = 0; c < K; ++c) {
                const double dist = M[(int64_t)cent[c] * m + t];
                if (dist < bd) { bd = dist; b = c; }
            }
            best[t] = b;
        }
        __syncthreads();
        for (int t = tid; t < m; t += blockDim.x) {                                           // :358-369, the row's sum over its own cluster
            const int c = best[t];
            double s = 0.0;
            if (c >= 0)
                for (int t1 = 0; t1 < m; ++t1)
                    if (best[t1] == c) s = __dadd_rn(s, M[(int64_t)t * m + t1]);
            sums[t] = s;
        }
        __syncthreads();
        for (int c = 0; c < K; ++c) {                                                         // :353-375
            double bv = 1.7976931348623157e308; int bi = -1;
            for (int t = tid; t < m; t += blockDim.x)
                if (best[t] == c && sums[t] < bv) { bv = sums[t]; bi = t; }
            s_v[tid] = bv; s_i[tid] = bi;
            __syncthreads();
            for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
                if (tid < o) {
                    const double ov = s_v[tid + o]; const int oi = s_i[tid + o];
                    if (oi >= 0 && (s_i[tid] < 0 || ov < s_v[tid] || (ov == s_v[tid] && oi < s_i[tid]))) { s_v[tid] = ov; s_i[tid] = oi; }
                }
                __syncthreads();
            }
            if (tid == 0) { cent[c] = s_i[0]; if (s_i[0] < 0) s_dead = 1; }
            __syncthreads();
        }
        if (s_dead) break;
    }
    if (tid == 0) dead_all[blockIdx.x] = s_dead;
}

}  // namespace fir

using namespace fir;

extern "C" int fir_kmedoids_select(const double* train_rows, const int32_t* train_labels, int64_t n, int32_t d, int32_t n_classes,
                                   int32_t num_clusters, int64_t* out_selected, int64_t* out_count) {
    if (!train_rows || !train_labels || !out_selected || !out_count || n <= 0 || d <= 0 || n_classes <= 0 || num_clusters < 1)
        return fail(FIR_ERR_BAD_ARG, "bad arguments");
    std::vector<int64_t> cls_begin((size_t)n_classes + 1, 0);
    for (int64_t t = 0; t < n; ++t) {
        const int32_t c = train_labels[t];
        if (c < 0 || c >= n_classes) return fail(FIR_ERR_BAD_ARG, "label out of range");
        if (t > 0 && c < train_labels[t - 1]) return fail(FIR_ERR_BAD_ARG, "training rows must be class-major");
        cls_begin[c + 1]++;
    }
    for (int c = 0; c < n_classes; ++c) cls_begin[c + 1] += cls_begin[c];
    // classes that need clustering, batched so that the pairwise matrices of one batch stay under 1 GiB
    std::vector<KmedClass> todo;
    std::vector<int> todo_class;
    for (int c = 0; c < n_classes; ++c) {
        const int64_t m = cls_begin[c + 1] - cls_begin[c];
        if (m > 46000) return fail(FIR_ERR_UNSUPPORTED, "class too large for the pairwise matrix");
        if (m > num_clusters) { todo.push_back(KmedClass{cls_begin[c], 0, (int32_t)m, 0}); todo_class.push_back(c); }
    }
    std::vector<std::vector<int32_t> > medoids((size_t)n_classes);
    if (!todo.empty()) {
        double* d_rows = nullptr; double* d_M = nullptr; double* d_sums = nullptr; int32_t* d_best = nullptr; int32_t* d_cent = nullptr; int32_t* d_dead = nullptr;
        KmedClass* d_cls = nullptr;
        auto release = [&]() { cudaFree(d_rows); cudaFree(d_M); cudaFree(d_sums); cudaFree(d_best); cudaFree(d_cent); cudaFree(d_dead); cudaFree(d_cls); };
        const int64_t budget = (int64_t)1 << 27;                       // doubles per batch of matrices (1 GiB)
        int64_t biggest = 0, all = 0;
        for (const KmedClass& k : todo) { biggest = std::max<int64_t>(biggest, (int64_t)k.m * k.m); all += (int64_t)k.m * k.m; }
        const int64_t m_cap = std::max(biggest, std::min(all, budget));
        cudaError_t e;
        if ((e = cudaMalloc(&d_rows, (size_t)n * d * 8)) != cudaSuccess || (e = cudaMalloc(&d_M, (size_t)m_cap * 8)) != cudaSuccess ||
            (e = cudaMalloc(&d_sums, (size_t)n * 8)) != cudaSuccess || (e = cudaMalloc(&d_best, (size_t)n * 4)) != cudaSuccess ||
            (e = cudaMalloc(&d_cent, todo.size() * (size_t)num_clusters * 4)) != cudaSuccess || (e = cudaMalloc(&d_dead, todo.size() * 4)) != cudaSuccess ||
            (e = cudaMalloc(&d_cls, todo.size() * sizeof(KmedClass))) != cudaSuccess ||
            (e = cudaMemcpy(d_rows, train_rows, (size_t)n * d * 8, cudaMemcpyHostToDevice)) != cudaSuccess) {
            release();
            return fail(e == cudaErrorMemoryAllocation ? FIR_ERR_OOM : FIR_ERR_CUDA, cudaGetErrorString(e));
        }
        std::vector<int32_t> cent(todo.size() * (size_t)num_clusters), dead(todo.size());
        for (size_t b0 = 0; b0 < todo.size();) {
            size_t b1 = b0; int64_t used = 0; int max_m = 0;
            while (b1 < todo.size() && (b1 == b0 || used + (int64_t)todo[b1].m * todo[b1].m <= m_cap)) {
                todo[b1].m_off = used; todo[b1].out_off = (int32_t)((b1 - b0) * num_clusters);
                used += (int64_t)todo[b1].m * todo[b1].m; max_m = std::max(max_m, todo[b1].m); ++b1;
            }
            const unsigned nb = (unsigned)(b1 - b0);
            e = cudaMemcpy(d_cls, todo.data() + b0, nb * sizeof(KmedClass), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) {
                dim3 grid((unsigned)std::min<int64_t>(ceil_div((int64_t)max_m * max_m, 256), 4096), nb);
                kmed_matrix_kernel<<<grid, 256>>>(d_rows, d, d_cls, d_M);
                kmed_iterate_kernel<<<nb, 256>>>(d_M, d_cls, num_clusters, 100, d_best, d_sums, d_cent, d_dead);
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaMemcpy(cent.data(), d_cent, nb * (size_t)num_clusters * 4, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(dead.data(), d_dead, nb * 4, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { release(); return fail(FIR_ERR_CUDA, cudaGetErrorString(e)); }
            for (unsigned i = 0; i < nb; ++i) {
                if (dead[i]) { release(); return fail(FIR_ERR_UNSUPPORTED, "a cluster ran empty (the reference then indexes with (size_t)-1)"); }
                medoids[todo_class[b0 + i]].assign(cent.begin() + (size_t)i * num_clusters, cent.begin() + (size_t)(i + 1) * num_clusters);
            }
            b0 = b1;
        }
        release();
    }
    int64_t total = 0;
    for (int c = 0; c < n_classes; ++c) {
        const int64_t lo = cls_begin[c], m = cls_begin[c + 1] - lo;
        if (m <= num_clusters) for (int64_t j = 0; j < m; ++j) out_selected[total++] = lo + j;                 // :381-386
        else for (int k = 0; k < num_clusters; ++k) out_selected[total++] = lo + medoids[c][k];               // :376-380
    }
    *out_count = total;
    return FIR_OK;
}
