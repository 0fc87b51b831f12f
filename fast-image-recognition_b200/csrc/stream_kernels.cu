// stream_kernels.cu — small-batch ("latency mode") path: up to 8 queries against the whole gallery in ONE pass over HBM.
//
// The 64x64 tile kernel (exact_kernels.cu) amortises each gallery row over 64 queries; with a handful of queries it
// would spend 63/64 of its arithmetic on padding.  Here a warp owns 32 gallery rows, stages them 64 dimensions at a
// time through a double-buffered cp.async ring, and lane c walks row c once for all NQ queries — the gallery is read
// exactly once (N·D·4 bytes, the HBM roofline of SURVEY.md §8(d) "single-query mode"), every distance is still the
// reference's sequential fp32 sum (feature_distance, qt_cpp/db_features.cpp:22-42; query = lhs as in ann.h:37).
// The Q x N distances land in a scratch array and are reduced by the kernels below: per-segment top-k (merged by
// merge_parts_kernel), per-class minimum, or the Parzen sums of PNN, all in gallery order.
#include "fir_common.cuh"
#include "handles.hpp"
#include <algorithm>

namespace fir {

__device__ __forceinline__ void s_cp_async16(void* smem, const void* gmem, int src_bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}

constexpr int SW = 4;          // warps per block
constexpr int SCH = 32;        // dims per slab (32: 9 KB per warp in flight ⇒ 5 blocks / 20 warps per SM)
constexpr int SLD = SCH + 4;   // row stride in floats (36): LDS.128 conflict-free across 8 consecutive rows

template <int METRIC, int NQ>
__global__ void __launch_bounds__(SW * 32) stream_distance_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ x, int ldx, int64_t n,
                                                                  int d_end, float* __restrict__ out, int64_t out_stride, int nq_live) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int dq_pad = (d_end + SCH - 1) / SCH * SCH;
    float* qs = reinterpret_cast<float*>(smem_raw);                                  // [NQ][dq_pad]
    float* tiles = qs + NQ * dq_pad;                                                 // [SW][2][32][SLD]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < NQ * dq_pad; i += SW * 32) {
        const int qi = i / dq_pad, c = i - qi * dq_pad;
        qs[i] = c < ldq ? q[(int64_t)qi * ldq + c] : 0.f;
    }
    __syncthreads();
    float* wt = tiles + (size_t)warp * 2 * 32 * SLD;
    const float log2c = METRIC == FIR_KL ? glibc_logf(2.0f) : 0.f;
    const int nslab = (d_end + SCH - 1) / SCH;
    const int64_t ngroups = (n + 31) / 32;
    for (int64_t grp = (int64_t)blockIdx.x * SW + warp; grp < ngroups; grp += (int64_t)gridDim.x * SW) {
        const int64_t r0 = grp * 32;
        auto stage = [&](int slab, int buf) {
            float* t = wt + buf * 32 * SLD;
            // 32 rows x (SCH/4) float4: lane handles float4 column (lane % CPR) of rows (lane / CPR) + RPI*i
            constexpr int CPR = SCH / 4, RPI = 32 / CPR;
            const int c4 = (lane % CPR) * 4, rr = lane / CPR;
            const int col = slab * SCH + c4;
#pragma unroll
            for (int i = 0; i < 32 / RPI; ++i) {
                const int r = rr + RPI * i;
                const int64_t row = r0 + r;
                const bool ok = row < n && col < ldx;
                s_cp_async16(&t[r * SLD + c4], x + (ok ? row : 0) * ldx + (ok ? col : 0), ok ? 16 : 0);
            }
            asm volatile("cp.async.commit_group;\n" ::);
        };
        float acc[NQ];
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) acc[qi] = 0.f;
        stage(0, 0);
        for (int sl = 0; sl < nslab; ++sl) {
            if (sl + 1 < nslab) { stage(sl + 1, (sl + 1) & 1); asm volatile("cp.async.wait_group 1;\n" ::); }
            else asm volatile("cp.async.wait_group 0;\n" ::);
            __syncwarp();
            const float* row = wt + (sl & 1) * 32 * SLD + lane * SLD;
            const int k0 = sl * SCH;
            const int kmax = min(SCH, d_end - k0);
            if constexpr (METRIC == FIR_KL) {
                // the query element of a step is warp-uniform here (every lane walks its own gallery row against the same query):
                // zero elements take the r·logf(2) shortcut (kl_step_uniform1).  One query at a time with the dimension loop
                // rolled: the live code is one step (two logf bodies), not NQ x 4 of them.
#pragma unroll
                for (int qi = 0; qi < NQ; ++qi) {
                    if (qi >= nq_live) break;                                  // uniform; spare query slots are never stored
                    float a = acc[qi];
                    const float* qrow = &qs[qi * dq_pad + k0];
#pragma unroll 1
                    for (int kk = 0; kk < kmax; ++kk) {
                        const float r = row[kk];
                        const bool light_ok = !__any_sync(0xffffffffu, r > 1.7014118e38f);
                        kl_step_uniform1(a, qrow[kk], r, log2c, light_ok);
                    }
                    acc[qi] = a;
                }
            } else {
            int kk = 0;
            for (; kk + 4 <= kmax; kk += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&row[kk]);
#pragma unroll
                for (int qi = 0; qi < NQ; ++qi) {
                    const float4 u = *reinterpret_cast<const float4*>(&qs[qi * dq_pad + k0 + kk]);
                    dist_step<METRIC>(acc[qi], u.x, v.x); dist_step<METRIC>(acc[qi], u.y, v.y);
                    dist_step<METRIC>(acc[qi], u.z, v.z); dist_step<METRIC>(acc[qi], u.w, v.w);
                }
            }
            for (; kk < kmax; ++kk)
#pragma unroll
                for (int qi = 0; qi < NQ; ++qi) dist_step<METRIC>(acc[qi], qs[qi * dq_pad + k0 + kk], row[kk]);
            }
            __syncwarp();
        }
        const int64_t myrow = r0 + lane;
        if (myrow < n) {
            const float den = (float)d_end;
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi)
                if (qi < nq_live) out[(int64_t)qi * out_stride + myrow] = __fdiv_rn(acc[qi], den);    // `out` holds nq_live rows, not NQ
        }
    }
}

template <int METRIC>
static int launch_stream_metric(int nq, const float* q, int ldq, const float* x, int ldx, int64_t n, int d_end, float* out, int64_t out_stride,
                                int n_sm, cudaStream_t s) {
    const int dq_pad = (d_end + SCH - 1) / SCH * SCH;
    const int NQ = nq <= 1 ? 1 : (nq <= 2 ? 2 : (nq <= 4 ? 4 : 8));
    const size_t smem = sizeof(float) * ((size_t)NQ * dq_pad + (size_t)SW * 2 * 32 * SLD);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(ceil_div(n, 32), SW), (int64_t)n_sm * 6));
    auto go = [&](auto kern) -> int {
        FIR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, SW * 32, smem, s>>>(q, ldq, x, ldx, n, d_end, out, out_stride, nq);
        FIR_CUDA_TRY(cudaGetLastError());
        return FIR_OK;
    };
    switch (NQ) {
        case 1: return go(stream_distance_kernel<METRIC, 1>);
        case 2: return go(stream_distance_kernel<METRIC, 2>);
        case 4: return go(stream_distance_kernel<METRIC, 4>);
        default: return go(stream_distance_kernel<METRIC, 8>);
    }
}

// q must hold 8 (NQ rounded up) rows; rows beyond nq may be anything (they are computed and dropped: `out` has nq rows)
int launch_stream_distances(int metric, int nq, const float* q, int ldq, const float* x, int ldx, int64_t n, int d_end, float* out,
                            int64_t out_stride, int n_sm, cudaStream_t s) {
    switch (metric) {
        case FIR_L2: return launch_stream_metric<FIR_L2>(nq, q, ldq, x, ldx, n, d_end, out, out_stride, n_sm, s);
        case FIR_CHI2: return launch_stream_metric<FIR_CHI2>(nq, q, ldq, x, ldx, n, d_end, out, out_stride, n_sm, s);
        case FIR_KL: return launch_stream_metric<FIR_KL>(nq, q, ldq, x, ldx, n, d_end, out, out_stride, n_sm, s);
    }
    return fail(FIR_ERR_BAD_ARG, "unknown metric");
}

// ---------------------------------------------------------------------------------------------------
// top-k of one query's distance array over one segment: per-thread sorted lists (indices ascend within a thread, so a
// strict '<' keeps the lower index on ties), then k rounds of a block-wide (dist, idx) arg-min over the list heads.
// ---------------------------------------------------------------------------------------------------
constexpr int SK_MAX = 16;

__global__ void __launch_bounds__(256) stream_topk_kernel(const float* __restrict__ dist, int64_t stride, int64_t n, int k, int nseg,
                                                          float* __restrict__ part_d, int32_t* __restrict__ part_i) {
    __shared__ float s_d[8];
    __shared__ int s_i[8];
    __shared__ int s_owner[8];
    __shared__ int s_win;
    const int seg = blockIdx.x, qi = blockIdx.y;
    const int64_t per = (n + nseg - 1) / nseg;
    const int64_t lo = seg * per, hi = min(n, lo + per);
    const float* dr = dist + (int64_t)qi * stride;
    float ld[SK_MAX]; int li[SK_MAX];
#pragma unroll
    for (int r = 0; r < SK_MAX; ++r) { ld[r] = 100000.0f; li[r] = -1; }      // nothing ≥ 100000 is accepted (ann.cpp:116)
    for (int64_t j = lo + threadIdx.x; j < hi; j += 256) {
        float d = dr[j]; int idx = (int)j;
        if (d < ld[SK_MAX - 1]) {
#pragma unroll
            for (int r = 0; r < SK_MAX; ++r)
                if (d < ld[r]) { float td = ld[r]; ld[r] = d; d = td; int ti = li[r]; li[r] = idx; idx = ti; }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int head = 0;
    for (int r = 0; r < k; ++r) {
        float hd = 100000.0f; int hi_ = -1;
        // head element of this thread's list (static indexing through a small unrolled select)
#pragma unroll
        for (int t = 0; t < SK_MAX; ++t) if (t == head) { hd = ld[t]; hi_ = li[t]; }
        float bd = hd; int bi = hi_ < 0 ? 0x7fffffff : hi_; int bo = threadIdx.x;
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int oo = __shfl_xor_sync(0xffffffffu, bo, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; bo = oo; }
        }
        if (lane == 0) { s_d[warp] = bd; s_i[warp] = bi; s_owner[warp] = bo; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float wd = s_d[0]; int wi = s_i[0], wo = s_owner[0];
            for (int w = 1; w < 8; ++w)
                if (s_d[w] < wd || (s_d[w] == wd && s_i[w] < wi)) { wd = s_d[w]; wi = s_i[w]; wo = s_owner[w]; }
            const int64_t o = ((int64_t)qi * nseg + seg) * k + r;
            const bool ok = wi != 0x7fffffff && wd < 100000.0f;
            part_d[o] = ok ? wd : 100000.0f;
            part_i[o] = ok ? wi : -1;
            s_win = ok ? wo : -1;
        }
        __syncthreads();
        if (s_win == (int)threadIdx.x) ++head;
        if (head >= SK_MAX) head = SK_MAX - 1;      // k <= SK_MAX, a list can never be popped more than k times
    }
}

// class reductions over a query's distance array: each thread folds a contiguous run of 64 rows in gallery order
// (run-length over the class-major labels) and flushes with one atomic per class change.
__global__ void __launch_bounds__(128) stream_class_kernel(const float* __restrict__ dist, int64_t stride, int64_t n, int nq,
                                                           const int32_t* __restrict__ labels, int n_classes, int mode, double two_var,
                                                           unsigned long long* __restrict__ cls_key, double* __restrict__ cls_score) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t lo = t * 64, hi = min(n, lo + 64);
    if (lo >= n) return;
    for (int qi = 0; qi < nq; ++qi) {
        const float* dr = dist + (int64_t)qi * stride;
        int cur_c = -1; unsigned long long cur_key = ~0ull; double cur_sum = 0.0;
        for (int64_t j = lo; j < hi; ++j) {
            const float d = dr[j];
            const int c = labels[j];
            if (mode == MODE_CLASSMIN) {
                if (!(d < 100000.0f)) continue;
                const unsigned long long key = ((unsigned long long)ordered_bits(d) << 32) | (uint32_t)j;
                if (c != cur_c) { if (cur_c >= 0) atomicMin(&cls_key[(int64_t)qi * n_classes + cur_c], cur_key); cur_c = c; cur_key = key; }
                else if (key < cur_key) cur_key = key;
            } else {
                const double e = exp(-(double)d / two_var);
                if (c != cur_c) { if (cur_c >= 0) atomicAdd(&cls_score[(int64_t)qi * n_classes + cur_c], cur_sum); cur_c = c; cur_sum = e; }
                else cur_sum += e;
            }
        }
        if (cur_c >= 0) {
            if (mode == MODE_CLASSMIN) atomicMin(&cls_key[(int64_t)qi * n_classes + cur_c], cur_key);
            else atomicAdd(&cls_score[(int64_t)qi * n_classes + cur_c], cur_sum);
        }
    }
}

int launch_stream_topk(const float* dist, int64_t stride, int64_t n, int nq, int k, int nseg, float* part_d, int32_t* part_i, cudaStream_t s) {
    if (k > SK_MAX) return fail(FIR_ERR_INTERNAL, "stream top-k supports k <= 16");
    stream_topk_kernel<<<dim3((unsigned)nseg, (unsigned)nq), 256, 0, s>>>(dist, stride, n, k, nseg, part_d, part_i);
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}

int launch_stream_class(const float* dist, int64_t stride, int64_t n, int nq, const int32_t* labels, int n_classes, int mode, double two_var,
                        unsigned long long* cls_key, double* cls_score, cudaStream_t s) {
    stream_class_kernel<<<(unsigned)ceil_div(ceil_div(n, 64), 128), 128, 0, s>>>(dist, stride, n, nq, labels, n_classes, mode, two_var, cls_key, cls_score);
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}

}  // namespace fir
