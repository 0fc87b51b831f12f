// approx_search.cu — batched chi-square / KL top-k as "fast approximate pass + exact rerank + certificate".
//
// The exact tile kernel pays ~16 (chi²: IEEE division) to ~90 (KL: two glibc logf emulated in fp64) instructions per
// element.  For top-k search that is unnecessary: a pass with fast intrinsics (MUFU rcp / lg2, FMA) gives every
// distance within a small, rigorously bounded error of the reference value; each (query, gallery split) keeps its R
// smallest approximate distances, the survivors are re-evaluated with the reference's own arithmetic
// (pair_distance_kernel) and the same prune / select / certificate kernels as the tensor path decide — with this
// path's error model — whether the answer is provably the reference's.  Uncertified queries are re-run exactly.
//   chi²: every term (l−r)²/(l+r) ≥ 0 is computed with ≤ 2 ulp of relative error, so
//         |approx − reference| ≤ rel·reference with rel = (2D+64)·2⁻²⁴ (both sums' accumulation + the term error).
//   KL  : |approx − reference| ≤ E = [2e-6 + 2·1.386·(D+8)·2⁻²⁴]·(‖q‖₁+‖x‖₁)/D  (lg2.approx and fast-division error on
//         each logarithm, times |l|; accumulation error of both sums bounded through Σ|term| ≤ 1.386·(‖q‖₁+‖x‖₁)).
//   KL, entropy form (non-negative features — what the loader's L1 normalisation of ReLU'd CNN features produces):
//         Σ l·ln(2l/s) + r·ln(2r/s) = [Σ l·ln l + ln2·Σ l] + [Σ r·ln r + ln2·Σ r] − ln2·Σ s·log2(s),  s = l + r, 0·ln 0 = 0.
//         The two brackets are per-row constants (fp64, computed once per gallery / per query batch), so a pair costs ONE
//         lg2.approx and one FMA per dimension instead of a reciprocal, two logarithms and six multiplications.  The sum is
//         accumulated per 32-dimension chunk and the chunk sums added up, which bounds the accumulation error by
//         (34 + D/32)·2⁻²⁴·Σ|s·log2 s| ≤ (34 + D/32)·2⁻²⁴·Λ·(‖q‖₁+‖x‖₁), Λ = max(1, log2(1 / smallest positive element)).
//         E·D = (‖q‖₁+‖x‖₁)·[1.7e-7 + 2⁻²⁴·(Λ + 3.5) + (34 + D/32)·2⁻²⁴·Λ + (D+8)·2⁻²⁴·1.386]  (lg2.approx, rounding of s, the
//         chunked accumulation, and the reference's own sequential fp32 sum around the exact value).
// PNN class scores are NOT served from here: exp(−d/2var) turns a 1e-6 relative distance error into a score error well
// above the 1e-5 bar, so fir_pnn_scores always uses the exact kernels.
#include "fir_common.cuh"
#include "handles.hpp"
#include <algorithm>

namespace fir {

__device__ __forceinline__ void a_cp_async16(void* smem, const void* gmem, int src_bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}

template <int METRIC>
__device__ __forceinline__ void approx_step(float& acc, float l, float r) {
    const float s = l + r;
    if (METRIC == FIR_CHI2) {
        // FADD, FSUB, FMUL, FSETP + FSEL, MUFU.RCP, FFMA: s <= 0 (the reference's `l + r > 0` guard) turns the reciprocal's
        // argument into +inf, i.e. the term into an exact 0; rcp.approx is within 1 ulp, the FMA adds the term unrounded.
        // (.ftz: the plain form expands into a denormal-handling sequence.  Sums below 1e-30 are dropped with the guard:
        //  such a term is at most its own sum, far below one ulp of any distance that is not itself exactly zero.)
        const float d = l - r;
        float rc;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(s > 1e-30f ? s : __int_as_float(0x7f800000)));
        acc = fmaf(d * d, rc, acc);
    } else {
        const float inv = __fdividef(2.f, s);
        const float tl = l * __logf(l * inv), tr = r * __logf(r * inv);
        acc += (s > 0.f && l > 0.f) ? tl : 0.f;
        acc += (s > 0.f && r > 0.f) ? tr : 0.f;
    }
}

constexpr int ATS = 64, ACH = 32, ALDT = ACH + 4, ALDD = ATS + 1;

struct ApproxParams {
    const float* q; int64_t nq; int ldq;
    const float* x; int64_t n; int ldx;
    int d; int R; int nsplit; int64_t tiles_per_split;
    float* cand_val; int32_t* cand_idx; float* slot_bound;      // [nq][nsplit][R], [nq][nsplit]
    const double* q_ent; const double* x_ent;                   // KL entropy form: Σ v·ln v + ln2·Σ v per row
};

// same tiling as exact_tile_kernel (64 x 64 tiles, 4x4 pairs per thread, 2-stage cp.async ring over 32-dim chunks)
template <int METRIC, bool ENT = false>
__global__ void __launch_bounds__(256) approx_tile_kernel(ApproxParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* qs = reinterpret_cast<float*>(smem_raw);
    float* xs = qs + 2 * ATS * ALDT;
    float* ds = xs + 2 * ATS * ALDT;
    float* tkd = ds + ATS * ALDD;                       // [ATS][R]
    int* tki = reinterpret_cast<int*>(tkd + ATS * p.R);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t q0 = (int64_t)blockIdx.x * ATS;
    const int64_t ntiles = (p.n + ATS - 1) / ATS;
    const int64_t t_lo = (int64_t)blockIdx.y * p.tiles_per_split, t_hi = min(ntiles, t_lo + p.tiles_per_split);
    const int nchunks = (p.d + ACH - 1) / ACH;
    const float inv_d = 1.0f / (float)p.d;
    const int lr0 = tid >> 3, lc = (tid & 7) * 4;
    const int R = p.R;
    for (int i = tid; i < ATS * R; i += 256) { tkd[i] = __int_as_float(0x7f800000); tki[i] = -1; }
    for (int64_t t = t_lo; t < t_hi; ++t) {
        const int64_t x0 = t * ATS;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        auto load_chunk = [&](int c, int st) {
            const int kk = c * ACH + lc;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = lr0 + 32 * h;
                const int64_t qi = q0 + r, xr = x0 + r;
                const bool qok = qi < p.nq, xok = xr < p.n;
                a_cp_async16(&qs[(st * ATS + r) * ALDT + lc], p.q + (qok ? qi : 0) * p.ldq + kk, qok ? 16 : 0);
                a_cp_async16(&xs[(st * ATS + r) * ALDT + lc], p.x + (xok ? xr : 0) * p.ldx + kk, xok ? 16 : 0);
            }
        };
        load_chunk(0, 0);
        asm volatile("cp.async.commit_group;\n" ::);
        for (int c = 0; c < nchunks; ++c) {
            const int st = c & 1;
            if (c + 1 < nchunks) load_chunk(c + 1, st ^ 1);
            asm volatile("cp.async.commit_group;\n" ::);
            asm volatile("cp.async.wait_group 1;\n" ::);
            __syncthreads();
            const float* qb = qs + st * ATS * ALDT;
            const float* xb = xs + st * ATS * ALDT;
            // rows are zero padded to a multiple of 32 dims: padded dims contribute exactly 0 to both divergences
            if (ENT) {
                // entropy form: u += s·log2(max(s, tiny)) per dimension (FADD, FMNMX, MUFU.LG2, FFMA); one partial sum per
                // 32-dimension chunk, the chunk sums are added up below (bounded accumulation error)
                float part[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) part[a][b] = 0.f;
#pragma unroll 2
                for (int k4 = 0; k4 < ACH / 4; ++k4) {
                    float4 qa[4], xa[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) qa[a] = *reinterpret_cast<const float4*>(&qb[(ty + 16 * a) * ALDT + k4 * 4]);
#pragma unroll
                    for (int b = 0; b < 4; ++b) xa[b] = *reinterpret_cast<const float4*>(&xb[(tx + 16 * b) * ALDT + k4 * 4]);
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const float s0 = qa[a].x + xa[b].x, s1 = qa[a].y + xa[b].y, s2 = qa[a].z + xa[b].z, s3 = qa[a].w + xa[b].w;
                            part[a][b] = fmaf(s0, __log2f(fmaxf(s0, 1e-37f)), part[a][b]);
                            part[a][b] = fmaf(s1, __log2f(fmaxf(s1, 1e-37f)), part[a][b]);
                            part[a][b] = fmaf(s2, __log2f(fmaxf(s2, 1e-37f)), part[a][b]);
                            part[a][b] = fmaf(s3, __log2f(fmaxf(s3, 1e-37f)), part[a][b]);
                        }
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] += part[a][b];
            } else {
#pragma unroll 2
            for (int k4 = 0; k4 < ACH / 4; ++k4) {
                float4 qa[4], xa[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) qa[a] = *reinterpret_cast<const float4*>(&qb[(ty + 16 * a) * ALDT + k4 * 4]);
#pragma unroll
                for (int b = 0; b < 4; ++b) xa[b] = *reinterpret_cast<const float4*>(&xb[(tx + 16 * b) * ALDT + k4 * 4]);
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        approx_step<METRIC>(acc[a][b], qa[a].x, xa[b].x);
                        approx_step<METRIC>(acc[a][b], qa[a].y, xa[b].y);
                        approx_step<METRIC>(acc[a][b], qa[a].z, xa[b].z);
                        approx_step<METRIC>(acc[a][b], qa[a].w, xa[b].w);
                    }
            }
            }
            __syncthreads();
        }
        if (ENT) {
            // distance·D = [entropy constant of the query] + [of the gallery row] − ln2·Σ s·log2 s, combined in fp64
            double cq[4], cx[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { const int64_t qi = q0 + ty + 16 * a; cq[a] = qi < p.nq ? p.q_ent[qi] : 0.0; }
#pragma unroll
            for (int b = 0; b < 4; ++b) { const int64_t xi = x0 + tx + 16 * b; cx[b] = xi < p.n ? p.x_ent[xi] : 0.0; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    ds[(ty + 16 * a) * ALDD + tx + 16 * b] = (float)((cq[a] + cx[b] - 0.6931471805599453 * (double)acc[a][b]) / (double)p.d);
        } else {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) ds[(ty + 16 * a) * ALDD + tx + 16 * b] = acc[a][b] * inv_d;
        }
        __syncthreads();
        if (tid < ATS && q0 + tid < p.nq) {
            const int jmax = (int)min((int64_t)ATS, p.n - x0);
            const float* drow = ds + tid * ALDD;
            float* ld = tkd + tid * R;
            int* li = tki + tid * R;
            float worst = ld[R - 1];
            for (int j = 0; j < jmax; ++j) {
                const float d = drow[j];
                if (d < worst) {
                    int pos = R - 1;
                    while (pos > 0 && d < ld[pos - 1]) { ld[pos] = ld[pos - 1]; li[pos] = li[pos - 1]; --pos; }
                    ld[pos] = d; li[pos] = (int)(x0 + j);
                    worst = ld[R - 1];
                }
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < ATS * R; i += 256) {
        const int r = i / R, s = i - r * R;
        const int64_t qi = q0 + r;
        if (qi < p.nq) {
            const int64_t o = (qi * p.nsplit + blockIdx.y) * R + s;
            p.cand_val[o] = tkd[i];
            p.cand_idx[o] = tki[i];
            if (s == R - 1) p.slot_bound[qi * p.nsplit + blockIdx.y] = tki[i] >= 0 ? tkd[i] : __int_as_float(0x7f800000);
        }
    }
}

// ‖row‖₁ per row (warp per row) and its maximum (float bits, non-negative ⇒ unsigned order).
// The KL error model (E = abs_coef·(‖q‖₁ + max‖x‖₁)) rests on 2l/(l+r) ≤ 2, i.e. on NON-NEGATIVE features; the reference's
// own guards (`l + r > 0`, `l > 0`, db_features.cpp:33-36) also admit mixed-sign rows (PCA'd / un-ReLU'd features), where
// log(2l/(l+r)) is unbounded as l + r → 0⁺.  A row with a negative element therefore reports ‖row‖₁ = +inf (its query can
// never be certified and is re-run exactly) and raises neg_flag (a gallery with negatives routes KL to the exact kernels).
// Also, for the entropy form of KL: ent[r] = Σ v·ln v + ln2·Σ v over the positive elements (fp64), the row's smallest
// positive element (minpos_row, per query) and the smallest over all rows (minpos_all, float bits, atomicMin).
__global__ void row_l1_kernel(const float* __restrict__ rows, int64_t n, int d, int ld, float* __restrict__ out, unsigned int* __restrict__ max_bits,
                              unsigned int* __restrict__ neg_flag, double* __restrict__ ent = nullptr, float* __restrict__ minpos_row = nullptr,
                              unsigned int* __restrict__ minpos_all = nullptr) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    double s = 0.0, h = 0.0;
    float mp = __int_as_float(0x7f800000);
    bool neg = false;
    for (int c = lane; c < d; c += 32) {
        const float v = rows[r * ld + c];
        s += fabs((double)v); neg = neg || v < 0.f;
        if (v > 0.f) { h += (double)v * log((double)v); mp = fminf(mp, v); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o); h += __shfl_xor_sync(0xffffffffu, h, o);
        mp = fminf(mp, __shfl_xor_sync(0xffffffffu, mp, o));
    }
    neg = __any_sync(0xffffffffu, neg);
    if (lane == 0) {
        const float f = neg ? __int_as_float(0x7f800000) : __double2float_ru(s);
        if (out) out[r] = f;
        if (max_bits && !neg) atomicMax(max_bits, __float_as_uint(f));
        if (neg_flag && neg) atomicOr(neg_flag, 1u);
        if (ent) ent[r] = h + 0.6931471805599453 * s;
        if (minpos_row) minpos_row[r] = mp;
        if (minpos_all) atomicMin(minpos_all, __float_as_uint(mp));       // positive floats: bit order = value order
    }
}

int approx_search_topk(fir_gallery* g, const float* queries, int64_t nq, int k, int memspace, int32_t* out_idx, float* out_dist) {
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const int metric = g->metric;
    cudaStream_t s = g->stream;
    if (!g->d_l1max) {
        FIR_CUDA_TRY(cudaMalloc(&g->d_l1max, 256));
        FIR_CUDA_TRY(cudaMemsetAsync(g->d_l1max, 0, 256, s));
        const unsigned int inf_bits = 0x7f800000u;            // [2]: smallest positive gallery element (entropy form of KL)
        FIR_CUDA_TRY(cudaMemcpyAsync(g->d_l1max + 2, &inf_bits, 4, cudaMemcpyHostToDevice, s));
        if (metric == FIR_KL && !g->kl_ent) FIR_CUDA_TRY(cudaMalloc(&g->kl_ent, 8 * (size_t)g->n));
        row_l1_kernel<<<(unsigned)ceil_div(g->n, 8), 256, 0, s>>>(g->rows, g->n, g->d, g->dp, nullptr, reinterpret_cast<unsigned int*>(g->d_l1max),
                                                                  reinterpret_cast<unsigned int*>(g->d_l1max) + 1, g->kl_ent, nullptr,
                                                                  reinterpret_cast<unsigned int*>(g->d_l1max) + 2);
        FIR_CUDA_TRY(cudaGetLastError());
        unsigned int neg = 0;                                   // once per gallery: does any gallery element carry a minus sign?
        FIR_CUDA_TRY(cudaMemcpyAsync(&neg, reinterpret_cast<unsigned int*>(g->d_l1max) + 1, 4, cudaMemcpyDeviceToHost, s));
        FIR_CUDA_TRY(cudaStreamSynchronize(s));
        g->has_negative = neg != 0;
    }
    if (metric == FIR_KL && g->has_negative) return kApproxDeclined;      // mixed-sign gallery: the KL bound does not hold — exact kernels
    const int R = k <= 4 ? 8 : (k <= 12 ? 16 : 32);
    const int64_t qblocks = ceil_div(nq, ATS), ntiles = ceil_div(g->n, ATS);
    const int nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(ceil_div((int64_t)g->n_sm * 4, qblocks), ntiles), 32));
    const int rt = nsplit * R;
    const int64_t kFbFast = 64;
    const int nsplit_fast = (int)std::max<int64_t>(1, std::min<int64_t>(1024, ceil_div(g->n, 64)));
    const int nsplit_slow = (int)std::max<int64_t>(1, std::min<int64_t>(16, ceil_div(g->n, 64 * 8)));
    const size_t fb_cells = std::max<size_t>((size_t)kFbFast * nsplit_fast, (size_t)nq * nsplit_slow) * k;
    size_t need = al(4 * (size_t)nq * g->dp) + 3 * al(4 * (size_t)nq * rt) + al(4 * (size_t)nq * nsplit) + 4 * al(4 * (size_t)nq) + al(8 * (size_t)nq) +
                  2 * al(4 * (size_t)nq * k) + 2 * al(fb_cells * 4) + 16384;
    FIR_TRY(g->ws.reserve(need));
    const float* dq = nullptr;
    if (memspace == FIR_DEVICE && g->d == g->dp) dq = queries;
    else {
        float* buf = (float*)g->ws.take(4 * (size_t)nq * g->dp);
        if (!buf) return fail(FIR_ERR_INTERNAL, "workspace underestimated (approx queries)");
        if (memspace == FIR_HOST) {
            if (g->d != g->dp) FIR_CUDA_TRY(cudaMemsetAsync(buf, 0, 4 * (size_t)nq * g->dp, s));
            FIR_CUDA_TRY(cudaMemcpy2DAsync(buf, 4 * (size_t)g->dp, queries, 4 * (size_t)g->d, 4 * (size_t)g->d, (size_t)nq, cudaMemcpyHostToDevice, s));
        } else FIR_TRY(launch_pad_rows(queries, nq, g->d, buf, g->dp, s));
        dq = buf;
    }
    float* cand_val = (float*)g->ws.take(4 * (size_t)nq * rt);
    int32_t* cand_idx = (int32_t*)g->ws.take(4 * (size_t)nq * rt);
    float* cand_exact = (float*)g->ws.take(4 * (size_t)nq * rt);
    float* slot_bound = (float*)g->ws.take(4 * (size_t)nq * nsplit);
    float* q_l1 = (float*)g->ws.take(4 * (size_t)nq);
    const bool entropy = metric == FIR_KL && g->kl_ent != nullptr;       // non-negative gallery (a mixed-sign one was declined above)
    double* q_ent = entropy ? (double*)g->ws.take(8 * (size_t)nq) : nullptr;
    float* q_minpos = entropy ? (float*)g->ws.take(4 * (size_t)nq) : nullptr;
    if (entropy && (!q_ent || !q_minpos)) return fail(FIR_ERR_INTERNAL, "workspace underestimated (approx path, entropy form)");
    int32_t* flagged = (int32_t*)g->ws.take(4 * (size_t)nq);
    float* od = out_dist; int32_t* oi = out_idx;
    if (memspace == FIR_HOST || !out_dist) od = (float*)g->ws.take(4 * (size_t)nq * k);
    if (memspace == FIR_HOST) oi = (int32_t*)g->ws.take(4 * (size_t)nq * k);
    float* part_d = (float*)g->ws.take(fb_cells * 4);
    int32_t* part_i = (int32_t*)g->ws.take(fb_cells * 4);
    if (!cand_val || !cand_idx || !cand_exact || !slot_bound || !q_l1 || !flagged || !od || !oi || !part_d || !part_i)
        return fail(FIR_ERR_INTERNAL, "workspace underestimated (approx path)");
    int32_t* n_flagged = reinterpret_cast<int32_t*>(g->d_l1max + 4);
    float* max_bound = g->d_l1max + 5;
    FIR_CUDA_TRY(cudaMemsetAsync(g->d_l1max + 4, 0, 8, s));
    row_l1_kernel<<<(unsigned)ceil_div(nq, 8), 256, 0, s>>>(dq, nq, g->d, g->dp, q_l1, nullptr, nullptr, q_ent, q_minpos, nullptr);   // +inf for a mixed-sign query ⇒ exact re-run

    ApproxParams p{};
    p.q = dq; p.nq = nq; p.ldq = g->dp; p.x = g->rows; p.n = g->n; p.ldx = g->dp; p.d = g->d; p.R = R; p.nsplit = nsplit;
    p.tiles_per_split = ceil_div(ntiles, nsplit);
    p.cand_val = cand_val; p.cand_idx = cand_idx; p.slot_bound = slot_bound;
    p.q_ent = q_ent; p.x_ent = g->kl_ent;
    FIR_CUDA_TRY(cudaMemsetAsync(cand_idx, 0xFF, 4 * (size_t)nq * rt, s));
    FIR_CUDA_TRY(cudaMemsetAsync(slot_bound, 0xFF, 4 * (size_t)nq * nsplit, s));
    const size_t smem = sizeof(float) * (size_t)(4 * ATS * ALDT + ATS * ALDD) + (size_t)ATS * R * 8;
    dim3 grid((unsigned)qblocks, (unsigned)nsplit);
    auto* ev = g->prof_begin(FIR_KERNEL_APPROX_TILES);
    if (metric == FIR_CHI2) {
        FIR_CUDA_TRY(cudaFuncSetAttribute(approx_tile_kernel<FIR_CHI2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        approx_tile_kernel<FIR_CHI2><<<grid, 256, smem, s>>>(p);
    } else if (entropy) {
        FIR_CUDA_TRY(cudaFuncSetAttribute(approx_tile_kernel<FIR_KL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        approx_tile_kernel<FIR_KL, true><<<grid, 256, smem, s>>>(p);
    } else {
        FIR_CUDA_TRY(cudaFuncSetAttribute(approx_tile_kernel<FIR_KL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        approx_tile_kernel<FIR_KL><<<grid, 256, smem, s>>>(p);
    }
    g->prof_end(ev);
    FIR_CUDA_TRY(cudaGetLastError());

    ErrModel em{};
    em.d = g->d; em.dist_scale = 1.0;
    if (metric == FIR_CHI2) { em.kind = 1; em.rel = (2.0 * g->d + 64.0) * 5.9604644775390625e-08; }
    else if (entropy) {
        const double u = 5.9604644775390625e-08;
        em.kind = 3; em.rel = 0.0;
        em.abs_coef = (1.7e-7 + 3.5 * u + (g->d + 8.0) * u * 1.386) / (double)g->d;      // the Λ-independent part
        em.lam_coef = (u + (34.0 + g->d / 32.0) * u) / (double)g->d;                     // × Λ = max(1, log2(1 / smallest positive element))
        em.q_l1 = q_l1; em.x_l1_max = g->d_l1max; em.q_minpos = q_minpos; em.x_minpos = g->d_l1max + 2;
    } else {
        em.kind = 2; em.rel = 0.0;
        em.abs_coef = (2e-6 + 2.0 * 1.386 * (g->d + 8.0) * 5.9604644775390625e-08) / (double)g->d;
        em.q_l1 = q_l1; em.x_l1_max = g->d_l1max;
    }
    FIR_TRY(launch_prune(cand_val, cand_idx, nq, rt, k, em, s));
    FIR_TRY(launch_pair_distances(metric, dq, nq, g->dp, g->rows, g->dp, g->n, g->d, cand_idx, rt, 0, cand_exact, s));
    FIR_TRY(launch_select(cand_exact, cand_idx, slot_bound, nq, nsplit, R, k, em, g->index_offset, od, oi, flagged, n_flagged, nullptr, max_bound, s));
    FIR_TRY(exact_topk_device(g, dq, nq, k, g->d, flagged, n_flagged, part_d, part_i, nsplit_fast, od, oi, 0, kFbFast));
    if (nq > kFbFast) FIR_TRY(exact_topk_device(g, dq, nq, k, g->d, flagged, n_flagged, part_d, part_i, nsplit_slow, od, oi, kFbFast, nq - kFbFast));
    g->stats.gpu_launches += 5;
    g->stats.path_used = FIR_PATH_APPROX;
    g->stats.n_candidates = rt;
    g->stats.n_fallback = -2;        // resolved lazily from d_l1max[4]
    if (memspace == FIR_HOST) {
        FIR_CUDA_TRY(cudaMemcpyAsync(out_idx, oi, 4 * (size_t)nq * k, cudaMemcpyDeviceToHost, s));
        if (out_dist) FIR_CUDA_TRY(cudaMemcpyAsync(out_dist, od, 4 * (size_t)nq * k, cudaMemcpyDeviceToHost, s));
        FIR_CUDA_TRY(cudaStreamSynchronize(s));
    }
    return FIR_OK;
}

}  // namespace fir
