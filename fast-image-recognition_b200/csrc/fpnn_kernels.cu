// fpnn_kernels.cu — FPNNClassifier of qt_cpp/classification.cpp (:618-791) on the GPU: the probabilistic neural network
// with a trigonometric orthogonal-series density estimate per (feature, class) instead of Parzen kernels over every
// training row, so that a query costs O(D·C·J) whatever the training-set size.
//
// Reference semantics:
//   normalize (:633-654)      val = clamp(features_scale·(x − avg)/std, ±0.5), 0 when std == 0
//   train (:658-695)          J = max(3, ceil(cbrt(N_train / C))); a[(fi·C + i)(2J+1)] = 0.5, then for every training row of
//                             class i and j < J:  a[..+2j+1] += cos(π(j+1)val)·(1/n_i)·(J−j)/(J(J+1)), a[..+2j+2] likewise with sin
//   predict_bf (:697-735)     cos/sin(π·val) once per feature, higher harmonics by the angle-addition recurrence (fp64),
//                             probab = a0 + Σ_j (a_c·cos_j + a_s·sin_j), outputs[i] += fasterlog2((float)probab) in fp32
//   predict_sequentional (:736-791)  the same in 32-feature chunks; after each chunk classes whose output is below
//                             max + fasterlog2(output_ratio)·max_fi are dropped (dropped classes are still re-counted on
//                             their stale outputs, :779-784); stop when one variant is left.
// Every product/sum is a separately rounded __dmul_rn/__dadd_rn (nvcc would otherwise contract them into FMAs), the fp32
// part restates fasterlog2 (:64-73) operation by operation.  cos()/sin() are CUDA's (≤ 2 ulp from glibc's), so the
// coefficients agree with the reference to ~1e-15 relative and labels agree unless two class outputs are that close.
#include "fir_common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

struct fir_fpnn {
    int device = 0;
    cudaStream_t stream = 0;
    int d = 0, n_classes = 0, J = 0;
    double scale = 1.0, pi = 0.0;
    double* a = nullptr;          // [d][C][2J+1]
    double* avg = nullptr;        // [d]
    double* sd = nullptr;         // [d]
    fir::Workspace ws;
};

namespace fir {

constexpr int FP_CHUNK = 32;      // = delta_features_count (classification.cpp:182)
constexpr int FP_MAXJ = 64;

__device__ __forceinline__ double fpnn_normalize(double x, double avg, double sd, double scale) {
    double val = (sd != 0.0) ? __ddiv_rn(__dmul_rn(scale, __dsub_rn(x, avg)), sd) : 0.0;      // :643
    if (val < -0.5) val = -0.5;
    else if (val > 0.5) val = 0.5;
    return val;
}

__device__ __forceinline__ float fasterlog2_dev(float x) {                                    // :64-73
    const uint32_t xi = __float_as_uint(x);
    const float mx = __uint_as_float((xi & 0x007FFFFFu) | (0x7eu << 23));
    float y = __uint2float_rn(xi);
    y = __fmul_rn(y, 1.1920928955078125e-07f);                                                // *= 1.0 / (1 << 23), exact
    return __fsub_rn(__fsub_rn(__fsub_rn(y, 124.22544637f), __fmul_rn(1.498030302f, mx)),
                     __fdiv_rn(1.72587999f, __fadd_rn(0.3520887068f, mx)));
}

// one thread per (feature fi, class i): walks the class's training rows in order (:681-691)
__global__ void fpnn_train_kernel(const double* __restrict__ train, const int32_t* __restrict__ cls_begin, int d, int n_classes, int J,
                                  const double* __restrict__ avg, const double* __restrict__ sd, double scale, double pi, double* __restrict__ a) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (int64_t)d * n_classes) return;
    const int fi = (int)(id % d), i = (int)(id / d);
    double* m = a + ((int64_t)fi * n_classes + i) * (2 * J + 1);
    const int64_t lo = cls_begin[i], hi = cls_begin[i + 1];
    const double cur_mult = __ddiv_rn(1.0, (double)(hi - lo));
    const double den = (double)((int64_t)J * (J + 1));
    m[0] = 0.5;
    for (int j = 0; j < 2 * J; ++j) m[1 + j] = 0.0;
    const double av = avg[fi], s = sd[fi];
    for (int64_t t = lo; t < hi; ++t) {
        const double val = fpnn_normalize(train[t * d + fi], av, s, scale);
        for (int j = 0; j < J; ++j) {
            const double ang = __dmul_rn(__dmul_rn(pi, (double)(j + 1)), val);
            const double w = (double)(J - j);
            m[2 * j + 1] = __dadd_rn(m[2 * j + 1], __ddiv_rn(__dmul_rn(__dmul_rn(cos(ang), cur_mult), w), den));   // :687
            m[2 * j + 2] = __dadd_rn(m[2 * j + 2], __ddiv_rn(__dmul_rn(__dmul_rn(sin(ang), cur_mult), w), den));   // :688
        }
    }
}

// one block per query, one thread per class (strided when C > blockDim)
__global__ void __launch_bounds__(256) fpnn_predict_kernel(const double* __restrict__ q, int64_t nq, int d, int n_classes, int J,
                                                           const double* __restrict__ a, const double* __restrict__ avg,
                                                           const double* __restrict__ sd, double scale, double pi, int sequential,
                                                           float output_delta, float* __restrict__ out_scratch,
                                                           unsigned char* __restrict__ check_scratch, int32_t* __restrict__ out_label) {
    extern __shared__ __align__(16) unsigned char fp_smem[];
    double* cs = reinterpret_cast<double*>(fp_smem);                 // [FP_CHUNK][J]
    double* sn = cs + FP_CHUNK * J;                                  // [FP_CHUNK][J]
    float* red_v = reinterpret_cast<float*>(sn + FP_CHUNK * J);      // [blockDim]
    int* red_i = reinterpret_cast<int*>(red_v + blockDim.x);         // [blockDim]
    __shared__ int s_best, s_variants, s_stop;
    const int64_t qi = blockIdx.x;
    const double* x = q + qi * d;
    float* outputs = out_scratch + qi * n_classes;
    unsigned char* check = check_scratch + qi * n_classes;
    const int tid = threadIdx.x;
    for (int i = tid; i < n_classes; i += blockDim.x) { outputs[i] = 0.f; check[i] = 1; }
    if (tid == 0) { s_best = -1; s_stop = 0; }
    __syncthreads();
    const int J1 = 2 * J + 1;
    for (int cur = 0; cur < d; cur += FP_CHUNK) {
        const int max_fi = min(cur + FP_CHUNK, d);
        if (tid < max_fi - cur) {                                    // :700-706 / :750-756
            const int fi = cur + tid;
            const double val = fpnn_normalize(x[fi], avg[fi], sd[fi], scale);
            const double ang = __dmul_rn(pi, val);
            const double c0 = cos(ang), s0 = sin(ang);
            double cj = c0, sj = s0;
            cs[tid * J] = c0; sn[tid * J] = s0;
            for (int j = 1; j < J; ++j) {
                const double cn = __dsub_rn(__dmul_rn(cj, c0), __dmul_rn(sj, s0));
                const double snn = __dadd_rn(__dmul_rn(cj, s0), __dmul_rn(sj, c0));
                cj = cn; sj = snn;
                cs[tid * J + j] = cj; sn[tid * J + j] = sj;
            }
        }
        __syncthreads();
        for (int i = tid; i < n_classes; i += blockDim.x) {
            if (!check[i]) continue;
            float o = outputs[i];
            for (int fi = cur; fi < max_fi; ++fi) {
                const double* m = a + ((int64_t)fi * n_classes + i) * J1;
                const double* cv = cs + (fi - cur) * J;
                const double* sv = sn + (fi - cur) * J;
                double probab = m[0];
                for (int j = 0; j < J; ++j)
                    probab = __dadd_rn(probab, __dadd_rn(__dmul_rn(m[2 * j + 1], cv[j]), __dmul_rn(m[2 * j + 2], sv[j])));   // :714 / :764
                o = __fadd_rn(o, fasterlog2_dev(__double2float_rn(probab)));                                               // :717 / :766
            }
            outputs[i] = o;
        }
        __syncthreads();
        if (!sequential && max_fi < d) continue;
        // arg-max over the classes still checked: strict '<' from -FLT_MAX, the first maximum wins (:721-728 / :770-776)
        float bv = -3.402823466e+38f; int bi = -1;
        for (int i = tid; i < n_classes; i += blockDim.x)
            if (check[i] && bv < outputs[i]) { bv = outputs[i]; bi = i; }
        red_v[tid] = bv; red_i[tid] = bi;
        __syncthreads();
        for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
            if (tid < o) {
                const float ov = red_v[tid + o]; const int oi = red_i[tid + o];
                if (oi >= 0 && (red_i[tid] < 0 || ov > red_v[tid] || (ov == red_v[tid] && oi < red_i[tid]))) { red_v[tid] = ov; red_i[tid] = oi; }
            }
            __syncthreads();
        }
        const float max_output = red_i[0] >= 0 ? red_v[0] : -3.402823466e+38f;
        if (tid == 0) { if (red_i[0] >= 0) s_best = red_i[0]; s_variants = 0; }
        __syncthreads();
        if (sequential) {
            const float thr = __fadd_rn(max_output, __fmul_rn(output_delta, (float)max_fi));  // :777
            int mine = 0;
            for (int i = tid; i < n_classes; i += blockDim.x) {                               // :779-784, every class, stale or not
                if (outputs[i] < thr) check[i] = 0;
                else ++mine;
            }
            if (mine) atomicAdd(&s_variants, mine);
            __syncthreads();
            if (tid == 0 && s_variants == 1) s_stop = 1;
            __syncthreads();
            if (s_stop) break;
        }
    }
    if (tid == 0) out_label[qi] = s_best;
}

}  // namespace fir

using namespace fir;

extern "C" {

int fir_fpnn_create(const double* train_rows, const int32_t* train_labels, int64_t n, int32_t d, int32_t n_classes, const double* avg,
                    const double* sd, double features_scale, fir_fpnn** out) {
    if (!out) return fail(FIR_ERR_BAD_ARG, "out is null");
    *out = nullptr;
    if (!train_rows || !train_labels || !avg || !sd || n <= 0 || d <= 0 || n_classes <= 0) return fail(FIR_ERR_BAD_ARG, "bad arguments");
    std::vector<int32_t> cls_begin((size_t)n_classes + 1, 0);
    for (int64_t t = 0; t < n; ++t) {
        const int32_t c = train_labels[t];
        if (c < 0 || c >= n_classes) return fail(FIR_ERR_BAD_ARG, "label out of range");
        if (t > 0 && c < train_labels[t - 1]) return fail(FIR_ERR_BAD_ARG, "training rows must be class-major (the order train() walks training_set)");
        cls_begin[c + 1]++;
    }
    for (int c = 0; c < n_classes; ++c) {
        if (cls_begin[c + 1] == 0) return fail(FIR_ERR_BAD_ARG, "a class without training rows (the reference divides by its size)");
        cls_begin[c + 1] += cls_begin[c];
    }
    int J = (int)std::ceil(std::pow(1.0 * (double)n / n_classes, 1.0 / 3));                   // :666
    if (J <= 3) J = 3;                                                                        // :671-673
    if (J > FP_MAXJ) return fail(FIR_ERR_UNSUPPORTED, "series length above 64");
    fir_fpnn* f = new fir_fpnn();
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { delete f; return fail(FIR_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e)); }
    f->device = dev; f->d = d; f->n_classes = n_classes; f->J = J; f->scale = features_scale; f->pi = std::atan(1.0) * 4;   // :656
    auto cleanup = [&](int code) { fir_fpnn_destroy(f); return code; };
    const size_t na = (size_t)d * n_classes * (2 * J + 1);
    double* d_train = nullptr; int32_t* d_begin = nullptr;
    auto bad = [&](cudaError_t err) { if (d_train) cudaFree(d_train); if (d_begin) cudaFree(d_begin); return cleanup(fail(err == cudaErrorMemoryAllocation ? FIR_ERR_OOM : FIR_ERR_CUDA, cudaGetErrorString(err))); };
    if ((e = cudaMalloc(&f->a, na * 8)) != cudaSuccess) return bad(e);
    if ((e = cudaMalloc(&f->avg, (size_t)d * 8)) != cudaSuccess) return bad(e);
    if ((e = cudaMalloc(&f->sd, (size_t)d * 8)) != cudaSuccess) return bad(e);
    if ((e = cudaMalloc(&d_train, (size_t)n * d * 8)) != cudaSuccess) return bad(e);
    if ((e = cudaMalloc(&d_begin, ((size_t)n_classes + 1) * 4)) != cudaSuccess) return bad(e);
    if ((e = cudaMemcpy(f->avg, avg, (size_t)d * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return bad(e);
    if ((e = cudaMemcpy(f->sd, sd, (size_t)d * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return bad(e);
    if ((e = cudaMemcpy(d_train, train_rows, (size_t)n * d * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return bad(e);
    if ((e = cudaMemcpy(d_begin, cls_begin.data(), ((size_t)n_classes + 1) * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return bad(e);
    fpnn_train_kernel<<<(unsigned)ceil_div((int64_t)d * n_classes, 128), 128>>>(d_train, d_begin, d, n_classes, J, f->avg, f->sd, features_scale, f->pi, f->a);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(d_train); cudaFree(d_begin); d_train = nullptr; d_begin = nullptr;
    if (e != cudaSuccess) return bad(e);
    *out = f;
    return FIR_OK;
}

int fir_fpnn_destroy(fir_fpnn* f) {
    if (!f) return FIR_OK;
    cudaSetDevice(f->device);
    if (f->a) cudaFree(f->a);
    if (f->avg) cudaFree(f->avg);
    if (f->sd) cudaFree(f->sd);
    f->ws.release();
    delete f;
    return FIR_OK;
}

int fir_fpnn_info(const fir_fpnn* f, int32_t* J, int64_t* n_coefficients) {
    if (!f) return fail(FIR_ERR_BAD_ARG, "handle is null");
    if (J) *J = f->J;
    if (n_coefficients) *n_coefficients = (int64_t)f->d * f->n_classes * (2 * f->J + 1);
    return FIR_OK;
}

int fir_fpnn_get_coefficients(const fir_fpnn* f, double* out_a) {
    if (!f || !out_a) return fail(FIR_ERR_BAD_ARG, "null argument");
    FIR_CUDA_TRY(cudaSetDevice(f->device));
    FIR_CUDA_TRY(cudaMemcpy(out_a, f->a, (size_t)f->d * f->n_classes * (2 * f->J + 1) * 8, cudaMemcpyDeviceToHost));
    return FIR_OK;
}

int fir_fpnn_predict(fir_fpnn* f, const double* queries, int64_t nq, int32_t sequential, float output_ratio, int32_t* out_label) {
    if (!f) return fail(FIR_ERR_BAD_ARG, "handle is null");
    if (nq < 0 || (nq > 0 && (!queries || !out_label))) return fail(FIR_ERR_BAD_ARG, "bad arguments");
    if (nq == 0) return FIR_OK;
    FIR_CUDA_TRY(cudaSetDevice(f->device));
    // fasterlog2(output_ratio), the constructor's output_delta (:621), with the same fp32 steps as the device code
    float output_delta;
    {
        uint32_t xi; std::memcpy(&xi, &output_ratio, 4);
        const uint32_t mi = (xi & 0x007FFFFFu) | (0x7eu << 23);
        float mx; std::memcpy(&mx, &mi, 4);
        volatile float y = (float)xi;
        y = (float)(y * (1.0 / (1 << 23)));
        volatile float t1 = y - 124.22544637f;
        volatile float t2 = 1.498030302f * mx;
        volatile float t3 = t1 - t2;
        volatile float t4 = 0.3520887068f + mx;
        volatile float t5 = 1.72587999f / t4;
        output_delta = t3 - t5;
    }
    const int C = f->n_classes, d = f->d;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const int64_t chunk = std::min<int64_t>(nq, 65536);
    FIR_TRY(f->ws.reserve(al((size_t)chunk * d * 8) + al((size_t)chunk * C * 4) + al((size_t)chunk * C) + al((size_t)chunk * 4) + 4096));
    double* dq = (double*)f->ws.take((size_t)chunk * d * 8);
    float* outs = (float*)f->ws.take((size_t)chunk * C * 4);
    unsigned char* check = (unsigned char*)f->ws.take((size_t)chunk * C);
    int32_t* lab = (int32_t*)f->ws.take((size_t)chunk * 4);
    if (!dq || !outs || !check || !lab) return fail(FIR_ERR_INTERNAL, "workspace underestimated (fpnn)");
    const int threads = C >= 256 ? 256 : (C > 128 ? 256 : (C > 64 ? 128 : (C > 32 ? 64 : 32)));
    const size_t smem = (size_t)2 * FP_CHUNK * f->J * 8 + (size_t)threads * 8;
    cudaStream_t s = f->stream;
    for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
        const int64_t m = std::min<int64_t>(chunk, nq - q0);
        FIR_CUDA_TRY(cudaMemcpyAsync(dq, queries + q0 * d, (size_t)m * d * 8, cudaMemcpyHostToDevice, s));
        fpnn_predict_kernel<<<(unsigned)m, threads, smem, s>>>(dq, m, d, C, f->J, f->a, f->avg, f->sd, f->scale, f->pi, sequential ? 1 : 0, output_delta,
                                                               outs, check, lab);
        FIR_CUDA_TRY(cudaGetLastError());
        FIR_CUDA_TRY(cudaMemcpyAsync(out_label + q0, lab, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
        FIR_CUDA_TRY(cudaStreamSynchronize(s));
    }
    return FIR_OK;
}

}  // extern "C"
