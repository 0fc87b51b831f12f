// exact_kernels.cu — bit-faithful CUDA-core kernels for the matching path.
//
// Every distance here reproduces feature_distance (qt_cpp/db_features.cpp:22-42) bit for bit: one
// thread owns a (query, gallery row) pair and accumulates over the dimensions in index order with
// separately rounded fp32 sub/mul/add(/div/logf) — parallelism comes from the Q x N pairs, never from
// splitting a sum.  The tile kernel fuses the reductions the reference performs right after the
// distance loop: argmin / top-k (ann.cpp:117-123, db_features.cpp:325-333), per-class minimum, and the
// Parzen sum of PNNClassifier::predict_bf (classification.cpp:213).
#include "fir_common.cuh"

namespace fir {

// ---------------------------------------------------------------------------------------------------
// cp.async helpers (LDGSTS): 16-byte global->shared copies, zero-filled when src_bytes == 0
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

#ifndef FIR_DIV_UNROLL
#define FIR_DIV_UNROLL 1
#endif
constexpr int kDivUnroll = FIR_DIV_UNROLL;   // k-loop unroll of the division / logf metrics (instruction-cache footprint)
constexpr int TS = kExactTile;        // 64
constexpr int CH = kExactChunk;       // 32
constexpr int LDT = CH + 4;           // smem row stride in floats: 36 ⇒ LDS.128 conflict-free for 8 consecutive rows
constexpr int LDD = TS + 1;

static size_t exact_smem_bytes(int k, int mode) {
    size_t b = sizeof(float) * (size_t)(4 * TS * LDT + TS * LDD) + sizeof(int) * TS;
    if (mode == MODE_TOPK) b += (size_t)TS * k * 8;
    return b;
}

// ---------------------------------------------------------------------------------------------------
// Tile kernel: block = 64 queries x (a split of the gallery, walked in 64-row tiles); 256 threads,
// thread (ty,tx) owns the 4x4 pairs {ty+16a} x {tx+16b}.  Dimensions are staged 32 at a time through a
// 2-stage cp.async ring; after the last chunk the 64x64 distances go to shared memory and one thread
// per query folds them, in gallery order, into its running reduction.
// ---------------------------------------------------------------------------------------------------
// THIN: the launch serves a device-side list of queries that is usually nearly empty (the certificate fallback of the
// tensor / approximate paths).  A warp owns query rows {2w, 2w+1} + 16a of the tile; with THIN it skips the arithmetic of
// the rows past the active count, so one stray query costs a pass over the gallery, not 64 queries' worth of FLOPs.
template <int METRIC, bool THIN>
__global__ void __launch_bounds__(256) exact_tile_kernel(ExactParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* qs = reinterpret_cast<float*>(smem_raw);   // [2][TS][LDT]
    float* xs = qs + 2 * TS * LDT;                    // [2][TS][LDT]
    float* ds = xs + 2 * TS * LDT;                    // [TS][LDD]
    int* ls = reinterpret_cast<int*>(ds + TS * LDD);  // [TS]
    float* tkd = reinterpret_cast<float*>(ls + TS);   // [TS][k]
    int* tki = reinterpret_cast<int*>(tkd + TS * p.k);

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // Pair ownership.  L2 / chi-square: thread (ty, tx) owns the 4 x 4 pairs {ty + 16a} x {tx + 16b}.  KL: warp w owns query rows
    // {w + 8a}, lane l gallery rows {l + 32b} — 8 x 2 pairs — so that the query element of a step is WARP-UNIFORM: for a zero query
    // element (half of ReLU-style features) the step is r·logf(2) for the lanes with r > 0, decided by a uniform branch instead of
    // two masked logf bodies (see kl_step_uniform).
    constexpr bool KLMAP = METRIC == FIR_KL;
    constexpr int NA = KLMAP ? 8 : 4, NB = KLMAP ? 2 : 4;
    const int qr0 = KLMAP ? (tid >> 5) : ty, qstep = KLMAP ? 8 : 16;       // query row of pair (a, .) = qr0 + qstep * a
    const int xr0 = KLMAP ? (tid & 31) : tx, xstep = KLMAP ? 32 : 16;      // gallery row of pair (., b) = xr0 + xstep * b
    int64_t n_active = p.n_active ? (int64_t)*p.n_active : p.nq;
    const int32_t* qmap_ = p.qmap ? p.qmap + p.active_offset : nullptr;          // chunked fallback: window of the active list
    if (p.n_active) { n_active = max((int64_t)0, n_active - p.active_offset); if (p.active_cap > 0) n_active = min(n_active, p.active_cap); }
    const int64_t q0 = (int64_t)blockIdx.x * TS;
    if (q0 >= n_active) return;
    const int rows_live = (int)min((int64_t)TS, n_active - q0);      // query rows of this tile that exist
    const int wrow = KLMAP ? (tid >> 5) : (tid >> 5) * 2;             // first query row of this warp (a = 0)
    const int64_t ntiles = (p.n + TS - 1) / TS;
    const int64_t t_lo = (int64_t)blockIdx.y * p.tiles_per_split;
    const int64_t t_hi = min(ntiles, t_lo + p.tiles_per_split);
    const int nchunks = (p.d_end + CH - 1) / CH;
    const float inv_den = (float)p.d_end;
    const float log2c = KLMAP ? glibc_logf(2.0f) : 0.f;              // the reference's own logf(2) (0x3f317218), computed by the same code

    // this thread's two (row, 16B column) slots of each 64x32 stage tile
    const int lr0 = tid >> 3, lc = (tid & 7) * 4;     // rows lr0 and lr0+32
    int64_t qrow[2];
    bool qok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int64_t qi = q0 + lr0 + 32 * h;
        qok[h] = qi < n_active;
        int64_t src = qok[h] ? (qmap_ ? (int64_t)qmap_[qi] : qi) : 0;
        qrow[h] = src * p.ldq;
    }

    if (p.mode == MODE_TOPK) {
        for (int i = tid; i < TS * p.k; i += 256) { tkd[i] = 100000.0f; tki[i] = -1; }
    }
    // per-query running state for the class reductions (only threads < TS use it)
    int cur_c = -1;
    unsigned long long cur_key = ~0ull;
    double cur_sum = 0.0;
    const int64_t my_q = q0 + tid;
    const int64_t my_q_out = (tid < TS && my_q < n_active) ? (qmap_ ? (int64_t)qmap_[my_q] : my_q) : -1;

    for (int64_t t = t_lo; t < t_hi; ++t) {
        const int64_t x0 = t * TS;
        float acc[KLMAP ? 1 : NA][KLMAP ? 1 : NB];
        if constexpr (KLMAP) {
            // (the previous tile's fold has finished reading ds: every thread passed the __syncthreads that ends the tile loop body)
#pragma unroll
            for (int a = 0; a < NA; ++a) { ds[(qr0 + 8 * a) * LDD + xr0] = 0.f; ds[(qr0 + 8 * a) * LDD + xr0 + 32] = 0.f; }
        } else {
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) acc[a][b] = 0.f;
        }

        auto load_chunk = [&](int c, int st) {
            const int kk = c * CH + lc;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int r = lr0 + 32 * h;
                cp_async16(&qs[(st * TS + r) * LDT + lc], p.q + qrow[h] + kk, qok[h] ? 16 : 0);
                int64_t xr = x0 + r;
                bool ok = xr < p.n;
                cp_async16(&xs[(st * TS + r) * LDT + lc], p.x + (ok ? xr : 0) * p.ldx + kk, ok ? 16 : 0);
            }
        };

        load_chunk(0, 0);
        cp_async_commit();
        if (tid < TS && p.mode != MODE_TOPK) ls[tid] = (x0 + tid < p.n) ? p.labels[x0 + tid] : -1;

        for (int c = 0; c < nchunks; ++c) {
            const int st = c & 1;
            if (c + 1 < nchunks) load_chunk(c + 1, st ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
            __syncthreads();
            const float* qb = qs + st * TS * LDT;
            const float* xb = xs + st * TS * LDT;
            const int kmax = min(CH, p.d_end - c * CH);
            if constexpr (!KLMAP) {
            if (kmax == CH) {
                // L2's 3-instruction step unrolls fully (24 KB of code); the division step is ~16 instructions, so its k-loop
                // stays rolled to fit the instruction cache (ncu: 28 % stall_no_inst when unrolled)
#pragma unroll (METRIC == FIR_L2 ? CH / 4 : kDivUnroll)
                for (int k4 = 0; k4 < CH / 4; ++k4) {
                    float4 qa[4], xa[4];
                    if (THIN && wrow >= rows_live) continue;                       // warp-uniform: none of this warp's rows exist
#pragma unroll
                    for (int a = 0; a < 4; ++a) qa[a] = *reinterpret_cast<const float4*>(&qb[(ty + 16 * a) * LDT + k4 * 4]);
#pragma unroll
                    for (int b = 0; b < 4; ++b) xa[b] = *reinterpret_cast<const float4*>(&xb[(tx + 16 * b) * LDT + k4 * 4]);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        if (THIN && wrow + 16 * a >= rows_live) continue;          // warp-uniform
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            dist_step<METRIC>(acc[a][b], qa[a].x, xa[b].x);
                            dist_step<METRIC>(acc[a][b], qa[a].y, xa[b].y);
                            dist_step<METRIC>(acc[a][b], qa[a].z, xa[b].z);
                            dist_step<METRIC>(acc[a][b], qa[a].w, xa[b].w);
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int kk = 0; kk < kmax; ++kk) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        if (THIN && wrow + 16 * a >= rows_live) continue;
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            dist_step<METRIC>(acc[a][b], qb[(ty + 16 * a) * LDT + kk], xb[(tx + 16 * b) * LDT + kk]);
                    }
                }
            }
            } else {
                // KL: the running sums live in the output tile `ds` between chunks, so the loops over the warp's query rows and over
                // the dimensions both stay ROLLED around two inlined steps (4 logf bodies, ~6 KB of code).  Fully unrolled the loop
                // body was 229 KB, with only the dimension loop rolled 57 KB — ncu: 36 % of the warp stalls were instruction-cache
                // misses (profiles/r2_ncu_exact_kl.txt); latency is hidden by the other warps, not by unrolling.
#pragma unroll 1
                for (int a = 0; a < NA; a += 2) {
                    if (THIN && wrow + 8 * a >= rows_live) break;                  // warp-uniform (rows ascend with a)
                    float* d0 = &ds[(qr0 + 8 * a) * LDD + xr0];
                    float* d1 = d0 + 8 * LDD;
                    float a00 = d0[0], a01 = d0[32], a10 = d1[0], a11 = d1[32];
                    const float* ql0 = &qb[(qr0 + 8 * a) * LDT];
                    const float* ql1 = ql0 + 8 * LDT;
                    const float* xr = &xb[xr0 * LDT];
#pragma unroll 1
                    for (int kk = 0; kk < kmax; ++kk) {
                        const float r0 = xr[kk], r1 = xr[32 * LDT + kk];
                        const bool light_ok = !__any_sync(0xffffffffu, r0 > 1.7014118e38f || r1 > 1.7014118e38f);   // 2r finite in every lane
                        kl_step_uniform(a00, a01, ql0[kk], r0, r1, log2c, light_ok);   // query elements: one address per warp, a broadcast
                        kl_step_uniform(a10, a11, ql1[kk], r0, r1, log2c, light_ok);
                    }
                    d0[0] = a00; d0[32] = a01; d1[0] = a10; d1[32] = a11;
                }
            }
            __syncthreads();
        }

        if constexpr (KLMAP) {
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) { float* o = &ds[(qr0 + 8 * a) * LDD + xr0 + 32 * b]; *o = __fdiv_rn(*o, inv_den); }   // this thread's own cells
        } else {
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) ds[(qr0 + qstep * a) * LDD + xr0 + xstep * b] = __fdiv_rn(acc[a][b], inv_den);   // db_features.cpp:40
        }
        __syncthreads();

        if (my_q_out >= 0) {
            const int jmax = (int)min((int64_t)TS, p.n - x0);
            const float* drow = ds + tid * LDD;
            if (p.mode == MODE_TOPK) {
                const int k = p.k;
                float* ld = tkd + tid * k;
                int* li = tki + tid * k;
                float worst = ld[k - 1];
                for (int j = 0; j < jmax; ++j) {
                    float d = drow[j];
                    if (d < worst) {          // strict '<': an equal distance never displaces an earlier index
                        int pos = k - 1;
                        while (pos > 0 && d < ld[pos - 1]) { ld[pos] = ld[pos - 1]; li[pos] = li[pos - 1]; --pos; }
                        ld[pos] = d;
                        li[pos] = (int)(x0 + j);
                        worst = ld[k - 1];
                    }
                }
            } else if (p.mode == MODE_CLASSMIN) {
                for (int j = 0; j < jmax; ++j) {
                    float d = drow[j];
                    if (!(d < 100000.0f)) continue;
                    int c = ls[j];
                    unsigned long long key = ((unsigned long long)ordered_bits(d) << 32) | (uint32_t)(x0 + j);
                    if (c != cur_c) {
                        if (cur_c >= 0) atomicMin(&p.cls_key[my_q_out * p.n_classes + cur_c], cur_key);
                        cur_c = c; cur_key = key;
                    } else if (key < cur_key) cur_key = key;
                }
            } else {
                for (int j = 0; j < jmax; ++j) {
                    int c = ls[j];
                    double e = exp(-(double)drow[j] / p.two_var);
                    if (c != cur_c) {
                        if (cur_c >= 0) atomicAdd(&p.cls_score[my_q_out * p.n_classes + cur_c], cur_sum);
                        cur_c = c; cur_sum = e;
                    } else cur_sum += e;
                }
            }
        }
        __syncthreads();
    }

    if (p.mode == MODE_TOPK) {
        for (int i = tid; i < TS * p.k; i += 256) {
            int r = i / p.k, s = i - r * p.k;
            int64_t qi = q0 + r;
            if (qi < n_active) {
                int64_t o = (qi * p.nsplit + blockIdx.y) * p.k + s;     // partial lists are indexed by ACTIVE position
                p.part_dist[o] = tkd[i];
                p.part_idx[o] = tki[i];
            }
        }
    } else if (my_q_out >= 0 && cur_c >= 0) {
        if (p.mode == MODE_CLASSMIN) atomicMin(&p.cls_key[my_q_out * p.n_classes + cur_c], cur_key);
        else atomicAdd(&p.cls_score[my_q_out * p.n_classes + cur_c], cur_sum);
    }
}

int launch_exact_tiles(int metric, const ExactParams& p, cudaStream_t s) {
    if (p.nq <= 0 || p.n <= 0) return FIR_OK;
    size_t smem = exact_smem_bytes(p.k, p.mode);
    if (smem > 200 * 1024) return fail(FIR_ERR_UNSUPPORTED, "k too large for the exact tile kernel");
    dim3 grid((unsigned)ceil_div(p.nq, TS), (unsigned)p.nsplit);
    auto go = [&](auto kern) -> int {
        FIR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 256, smem, s>>>(p);
        FIR_CUDA_TRY(cudaGetLastError());
        return FIR_OK;
    };
    const bool thin = p.n_active != nullptr;      // device-side query list: the certificate fallback, usually (nearly) empty
    switch (metric) {
        case FIR_L2: return thin ? go(exact_tile_kernel<FIR_L2, true>) : go(exact_tile_kernel<FIR_L2, false>);
        case FIR_CHI2: return thin ? go(exact_tile_kernel<FIR_CHI2, true>) : go(exact_tile_kernel<FIR_CHI2, false>);
        case FIR_KL: return thin ? go(exact_tile_kernel<FIR_KL, true>) : go(exact_tile_kernel<FIR_KL, false>);
    }
    return fail(FIR_ERR_BAD_ARG, "unknown metric");
}

// ---------------------------------------------------------------------------------------------------
// Lexicographic (dist, idx) merge of n_parts sorted lists per query.  idx < 0 marks an empty slot.
// Used for the gallery splits of one GPU and for the all-gathered per-GPU lists (fir_merge_topk).
// ---------------------------------------------------------------------------------------------------
// One warp per query.  Every part's current head is kept as a packed 64-bit key (ordered distance bits << 32 | index;
// ~0 = exhausted) in shared memory, lane l looking after parts l, l+32, ...; a round is a lane-local scan of those keys, a
// five-step shuffle tournament, and one global load by the lane whose part won.  The key order is exactly the
// lexicographic (dist, idx) order of the reference's strict '<' scan in index order.
constexpr int MERGE_WARPS = 4;
__device__ __forceinline__ unsigned long long merge_key(const float* __restrict__ pd, const int32_t* __restrict__ pi, int64_t o) {
    const int32_t ci = pi[o];
    return ci < 0 ? ~0ull : (((unsigned long long)ordered_bits(pd[o]) << 32) | (uint32_t)ci);
}
__global__ void __launch_bounds__(MERGE_WARPS * 32) merge_parts_kernel(const float* __restrict__ pd, const int32_t* __restrict__ pi, int n_parts,
                                   int64_t part_stride, int64_t q_stride, int64_t nq, int k, int64_t index_offset,
                                   const int32_t* qmap, const int32_t* n_active, float* od, int32_t* oi, int64_t active_offset,
                                   int64_t active_cap) {
    extern __shared__ __align__(16) unsigned char merge_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int np32 = (n_parts + 31) & ~31;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(merge_smem) + (size_t)warp * np32;
    unsigned short* head = reinterpret_cast<unsigned short*>(reinterpret_cast<unsigned long long*>(merge_smem) + (size_t)MERGE_WARPS * np32) + (size_t)warp * np32;
    const int64_t i = (int64_t)blockIdx.x * MERGE_WARPS + warp;
    int64_t na = n_active ? (int64_t)*n_active : nq;
    if (n_active) { na = max((int64_t)0, na - active_offset); if (active_cap > 0) na = min(na, active_cap); }
    if (i >= na) return;
    const int64_t q = qmap ? (int64_t)qmap[active_offset + i] : i;
    for (int s = lane; s < np32; s += 32) {                     // parts are indexed by active position, outputs by query
        keys[s] = s < n_parts ? merge_key(pd, pi, (int64_t)s * part_stride + i * q_stride) : ~0ull;
        head[s] = 0;
    }
    __syncwarp();
    for (int r = 0; r < k; ++r) {
        unsigned long long best = ~0ull; int bs = -1;
        for (int s = lane; s < np32; s += 32) { const unsigned long long v = keys[s]; if (v < best) { best = v; bs = s; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int os = __shfl_xor_sync(0xffffffffu, bs, o);
            if (ov < best) { best = ov; bs = os; }              // keys of live entries are distinct (an index appears once)
        }
        if (best == ~0ull) {                                    // every part exhausted: pad the rest
            for (int t = r + lane; t < k; t += 32) { od[q * k + t] = 0.f; oi[q * k + t] = -1; }
            break;
        }
        if ((bs & 31) == lane) {
            od[q * k + r] = from_ordered_bits((uint32_t)(best >> 32));
            oi[q * k + r] = (int32_t)((int64_t)(int32_t)(uint32_t)best + index_offset);
            const int h = ++head[bs];
            keys[bs] = h < k ? merge_key(pd, pi, (int64_t)bs * part_stride + i * q_stride + h) : ~0ull;
        }
        __syncwarp();
    }
}

int launch_merge_parts(const float* pd, const int32_t* pi, int n_parts, int64_t part_stride, int64_t q_stride,
                       int64_t nq, int k, int64_t index_offset, const int32_t* qmap, const int32_t* n_active,
                       float* od, int32_t* oi, cudaStream_t s, int64_t active_offset, int64_t active_cap) {
    if (nq <= 0) return FIR_OK;
    if (n_parts > 1024) return fail(FIR_ERR_UNSUPPORTED, "more than 1024 parts to merge");
    const size_t smem = (size_t)MERGE_WARPS * ((n_parts + 31) & ~31) * (8 + 2);
    merge_parts_kernel<<<(unsigned)ceil_div(nq, MERGE_WARPS), MERGE_WARPS * 32, smem, s>>>(pd, pi, n_parts, part_stride, q_stride, nq, k, index_offset,
                                                                   qmap, n_active, od, oi, active_offset, active_cap);
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}

__global__ void fill_u64_kernel(unsigned long long* p, int64_t n, unsigned long long v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
int launch_fill_u64(unsigned long long* p, int64_t n, unsigned long long v, cudaStream_t s) {
    if (n <= 0) return FIR_OK;
    fill_u64_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(p, n, v);
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}

__global__ void classmin_finalize_kernel(const unsigned long long* keys, int64_t n, int64_t index_offset, float* omin, int32_t* oarg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long key = keys[i];
    if (key == ~0ull) { omin[i] = 100000.0f; oarg[i] = -1; }
    else { omin[i] = from_ordered_bits((uint32_t)(key >> 32)); oarg[i] = (int32_t)((uint32_t)key + index_offset); }
}
int launch_classmin_finalize(const unsigned long long* keys, int64_t n, int64_t index_offset, float* omin, int32_t* oarg, cudaStream_t s) {
    if (n <= 0) return FIR_OK;
    classmin_finalize_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(keys, n, index_offset, omin, oarg);
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}

// scores /= n_total (classification.cpp:215), argmax with strict '<' from -DBL_MAX (:217-225)
__global__ void pnn_finalize_kernel(double* scores, int64_t nq, int n_classes, double n_total, int32_t* olabel) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    double mx = -1.7976931348623157e308; int best = -1;
    double* s = scores + q * n_classes;
    for (int c = 0; c < n_classes; ++c) {
        double v = s[c] / n_total;
        s[c] = v;
        if (mx < v) { mx = v; best = c; }
    }
    if (olabel) olabel[q] = best;
}
int launch_pnn_finalize(double* scores, int64_t nq, int n_classes, double n_total, int32_t* olabel, cudaStream_t s) {
    if (nq <= 0) return FIR_OK;
    pnn_finalize_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, s>>>(scores, nq, n_classes, n_total, olabel);
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}

// ---------------------------------------------------------------------------------------------------
// Exact distances for explicit (query, gallery row) pairs — the fp32 rerank of the tensor path and the
// candidate evaluation of directed enumeration.  One warp = one query x 32 candidates: the 32 candidate
// rows are staged 128 dims at a time into shared memory with coalesced 512-byte loads, then lane c
// walks row c sequentially.
// ---------------------------------------------------------------------------------------------------
constexpr int PW = 4;            // warps per block
constexpr int PCH = 128;         // dims per stage
constexpr int PLD = PCH + 4;     // 132 ⇒ LDS.128 conflict-free across 8 consecutive rows

template <int METRIC>
__global__ void __launch_bounds__(PW * 32) pair_distance_kernel(const float* __restrict__ q, int64_t nq, int ldq,
                                                                const float* __restrict__ x, int ldx, int64_t n, int d_end,
                                                                const int32_t* __restrict__ cand, int r, int gallery_is_lhs,
                                                                float* __restrict__ out, const int32_t* __restrict__ qsel,
                                                                const int32_t* __restrict__ qmap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = reinterpret_cast<float*>(smem_raw) + (size_t)warp * (32 * PLD + PCH);
    float* qv = tile + 32 * PLD;
    const int groups = (r + 31) / 32;
    const int64_t item = (int64_t)blockIdx.x * PW + warp;
    if (item >= nq * groups) return;
    const int64_t qi = item / groups;
    const int g = (int)(item - qi * groups);
    const int slot = g * 32 + lane;
    // cand == nullptr: candidate list is the identity (every gallery row); qsel: the query IS gallery row qsel[qi]
    int32_t my = (slot < r) ? (cand ? cand[qi * (int64_t)r + slot] : slot) : -1;
    if (my >= n) my = -1;
    if (__all_sync(0xffffffffu, my < 0)) {
        if (slot < r) out[qi * (int64_t)r + slot] = __int_as_float(0x7f800000);
        return;
    }
    // qmap: query qi of this launch is row qmap[qi] of the query matrix (active-list addressing)
    const float* qrow = qsel ? (x + (int64_t)qsel[qi] * ldx) : (q + (qmap ? (int64_t)qmap[qi] : qi) * ldq);
    const int ldq_eff = qsel ? ldx : ldq;
    float acc = 0.f;
    const int ld_lim = ldx;   // rows are zero padded up to ldx
    for (int c0 = 0; c0 < d_end; c0 += PCH) {
        const int col = c0 + lane * 4;
        // stage this 128-dim slab of every valid candidate row straight into shared memory with cp.async: all row
        // fetches of the slab are in flight together (one DRAM round trip per slab instead of one per row)
        for (int c = 0; c < 32; ++c) {
            const int32_t ci = __shfl_sync(0xffffffffu, my, c);
            if (ci < 0) { *reinterpret_cast<float4*>(&tile[c * PLD + lane * 4]) = make_float4(0.f, 0.f, 0.f, 0.f); continue; }   // warp-uniform
            const bool ok = col < ld_lim;
            cp_async16(&tile[c * PLD + lane * 4], x + (int64_t)ci * ldx + (ok ? col : 0), ok ? 16 : 0);
        }
        {
            const bool ok = col < ldq_eff;
            cp_async16(&qv[lane * 4], qrow + (ok ? col : 0), ok ? 16 : 0);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        const int kmax = min(PCH, d_end - c0);
        const float* row = tile + lane * PLD;
        int kk = 0;
        if (gallery_is_lhs) {
            for (; kk + 4 <= kmax; kk += 4) {
                float4 a = *reinterpret_cast<const float4*>(&row[kk]);
                float4 b = *reinterpret_cast<const float4*>(&qv[kk]);
                dist_step<METRIC>(acc, a.x, b.x); dist_step<METRIC>(acc, a.y, b.y);
                dist_step<METRIC>(acc, a.z, b.z); dist_step<METRIC>(acc, a.w, b.w);
            }
            for (; kk < kmax; ++kk) dist_step<METRIC>(acc, row[kk], qv[kk]);
        } else {
            for (; kk + 4 <= kmax; kk += 4) {
                float4 a = *reinterpret_cast<const float4*>(&row[kk]);
                float4 b = *reinterpret_cast<const float4*>(&qv[kk]);
                dist_step<METRIC>(acc, b.x, a.x); dist_step<METRIC>(acc, b.y, a.y);
                dist_step<METRIC>(acc, b.z, a.z); dist_step<METRIC>(acc, b.w, a.w);
            }
            for (; kk < kmax; ++kk) dist_step<METRIC>(acc, qv[kk], row[kk]);
        }
        __syncwarp();
    }
    if (slot < r) out[qi * (int64_t)r + slot] = (my >= 0) ? __fdiv_rn(acc, (float)d_end) : __int_as_float(0x7f800000);
}

int launch_pair_distances(int metric, const float* q, int64_t nq, int ldq, const float* x, int ldx, int64_t n, int d_end,
                          const int32_t* cand, int r, int gallery_is_lhs, float* out, cudaStream_t s, const int32_t* qsel,
                          const int32_t* qmap) {
    if (nq <= 0 || r <= 0) return FIR_OK;
    const int groups = (r + 31) / 32;
    size_t smem = sizeof(float) * (size_t)PW * (32 * PLD + PCH);
    unsigned grid = (unsigned)ceil_div(nq * groups, PW);
    auto go = [&](auto kern) -> int {
        FIR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, PW * 32, smem, s>>>(q, nq, ldq, x, ldx, n, d_end, cand, r, gallery_is_lhs, out, qsel, qmap);
        FIR_CUDA_TRY(cudaGetLastError());
        return FIR_OK;
    };
    switch (metric) {
        case FIR_L2: return go(pair_distance_kernel<FIR_L2>);
        case FIR_CHI2: return go(pair_distance_kernel<FIR_CHI2>);
        case FIR_KL: return go(pair_distance_kernel<FIR_KL>);
    }
    return fail(FIR_ERR_BAD_ARG, "unknown metric");
}

// Same arithmetic over an explicit list of (query, candidate) cells gathered from ALL queries (the survivors of the prune
// are few per query — about k — so one warp per query would run nearly empty).  One lane = one pair; both rows of every
// pair are staged 64 dims at a time with cp.async (two rows per instruction), then each lane walks its pair sequentially.
constexpr int LCH = 64;          // dims per stage
constexpr int LLD = LCH + 4;     // 68 ⇒ LDS.128 conflict-free across 8 consecutive rows

template <int METRIC>
__global__ void __launch_bounds__(PW * 32) pair_list_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ x, int ldx, int d_end,
                                                            const uint32_t* __restrict__ cells, const int32_t* __restrict__ count, int rt,
                                                            const int32_t* __restrict__ cand_idx, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total = *count;
    float* xs = reinterpret_cast<float*>(smem_raw) + (size_t)warp * (2 * 32 * LLD);
    float* qs = xs + 32 * LLD;
    // warps stride over the list: the grid is sized for the machine, not for the worst-case list (at C5 the list could hold
    // 77M cells and holds 1.5M — 600k blocks that exit at once cost more than the rerank itself)
    for (int64_t base = ((int64_t)blockIdx.x * PW + warp) * 32; base < total; base += (int64_t)gridDim.x * PW * 32) {
    const bool valid = base + lane < total;
    const uint32_t cell = valid ? cells[base + lane] : 0u;
    const int64_t my_q = valid ? (int64_t)(cell / (uint32_t)rt) : -1;
    const int64_t my_x = valid ? (int64_t)cand_idx[cell] : -1;
    const int sub = lane >> 4, l16 = lane & 15;
    float acc = 0.f;
    for (int c0 = 0; c0 < d_end; c0 += LCH) {
        const int col = c0 + l16 * 4;
#pragma unroll 4
        for (int c = 0; c < 16; ++c) {
            const int r = 2 * c + sub;
            const int64_t xr = __shfl_sync(0xffffffffu, my_x, r);
            const int64_t qr = __shfl_sync(0xffffffffu, my_q, r);
            if (xr < 0) continue;
            const bool okx = col < ldx, okq = col < ldq;
            cp_async16(&xs[r * LLD + l16 * 4], x + xr * ldx + (okx ? col : 0), okx ? 16 : 0);
            cp_async16(&qs[r * LLD + l16 * 4], q + qr * ldq + (okq ? col : 0), okq ? 16 : 0);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        if (valid) {
            const int kmax = min(LCH, d_end - c0);
            const float* xrow = xs + lane * LLD;
            const float* qrow = qs + lane * LLD;
            int kk = 0;
            for (; kk + 4 <= kmax; kk += 4) {
                const float4 a = *reinterpret_cast<const float4*>(&xrow[kk]);
                const float4 b = *reinterpret_cast<const float4*>(&qrow[kk]);
                dist_step<METRIC>(acc, b.x, a.x); dist_step<METRIC>(acc, b.y, a.y);            // lhs = query
                dist_step<METRIC>(acc, b.z, a.z); dist_step<METRIC>(acc, b.w, a.w);
            }
            for (; kk < kmax; ++kk) dist_step<METRIC>(acc, qrow[kk], xrow[kk]);
        }
        __syncwarp();
    }
    if (valid) out[cell] = __fdiv_rn(acc, (float)d_end);
    __syncwarp();
    }
}

int launch_pair_list(int metric, const float* q, int ldq, const float* x, int ldx, int d_end, const uint32_t* cells, const int32_t* count,
                     int64_t max_cells, int rt, const int32_t* cand_idx, float* out, cudaStream_t s) {
    if (max_cells <= 0) return FIR_OK;
    const size_t smem = sizeof(float) * (size_t)PW * (2 * 32 * LLD);
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(max_cells, (int64_t)PW * 32), 148 * 16);   // warps stride over the device-side count
    auto go = [&](auto kern) -> int {
        FIR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, PW * 32, smem, s>>>(q, ldq, x, ldx, d_end, cells, count, rt, cand_idx, out);
        FIR_CUDA_TRY(cudaGetLastError());
        return FIR_OK;
    };
    switch (metric) {
        case FIR_L2: return go(pair_list_kernel<FIR_L2>);
        case FIR_CHI2: return go(pair_list_kernel<FIR_CHI2>);
        case FIR_KL: return go(pair_list_kernel<FIR_KL>);
    }
    return fail(FIR_ERR_BAD_ARG, "unknown metric");
}

// ---------------------------------------------------------------------------------------------------
// Loader normalisation (qt_cpp/db_features.cpp:79-101), one thread per row, sequential fp32 sums.
// ---------------------------------------------------------------------------------------------------
// One warp per 32 rows.  Pass 1: 64-dim slabs are staged coalesced into shared memory and lane r walks row r in index
// order (the reference's sequential fp32 sum).  Pass 2: every element is re-read coalesced, zeroed / divided, written back.
constexpr int NCH = 64, NLD = NCH + 1;
__global__ void __launch_bounds__(128) normalize_rows_kernel(float* rows, int64_t n, int d, int ld, int metric) {
    __shared__ float tile[4][32 * NLD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* t = tile[warp];
    const int64_t r0 = ((int64_t)blockIdx.x * 4 + warp) * 32;
    if (r0 >= n) return;
    float sum = 0.f;
    for (int c0 = 0; c0 < d; c0 += NCH) {
        for (int r = 0; r < 32; ++r) {
#pragma unroll
            for (int h = 0; h < NCH / 32; ++h) {
                const int c = c0 + lane + 32 * h;
                t[r * NLD + lane + 32 * h] = (r0 + r < n && c < d) ? rows[(r0 + r) * ld + c] : 0.f;
            }
        }
        __syncwarp();
        const int kmax = min(NCH, d - c0);
        for (int kk = 0; kk < kmax; ++kk) {
            float v = t[lane * NLD + kk];
            if ((double)fabsf(v) < 0.0001) v = 0.f;                               // :85-86 (float |x| against a double literal)
            sum = (metric == FIR_L2 || metric == FIR_NORM_VIDEO_SUMSQ) ? __fadd_rn(sum, __fmul_rn(v, v)) : __fadd_rn(sum, v);   // :91 / :93; video.cpp:77
        }
        __syncwarp();
    }
    if (metric == FIR_L2) sum = __fsqrt_rn(sum);                                  // :98
    for (int r = 0; r < 32 && r0 + r < n; ++r) {
        const float sr = __shfl_sync(0xffffffffu, sum, r);
        float* f = rows + (r0 + r) * ld;
        for (int c = lane; c < d; c += 32) {
            float v = f[c];
            if ((double)fabsf(v) < 0.0001) v = 0.f;
            f[c] = __fdiv_rn(v, sr);                                              // :100-101
        }
    }
}
int launch_normalize_rows(float* rows, int64_t n, int d, int ld, int metric, cudaStream_t s) {
    if (n <= 0) return FIR_OK;
    normalize_rows_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, s>>>(rows, n, d, ld, metric);
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}

// rows [n][d] → zero-padded [n][ld]
__global__ void pad_rows_kernel(const float* __restrict__ src, int64_t n, int d, float* __restrict__ dst, int ld) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = n * ld;
    if (i >= total) return;
    int64_t r = i / ld;
    int c = (int)(i - r * ld);
    dst[i] = (c < d) ? src[r * d + c] : 0.f;
}
int launch_pad_rows(const float* src, int64_t n, int d, float* dst, int ld, cudaStream_t s) {
    if (n <= 0) return FIR_OK;
    pad_rows_kernel<<<(unsigned)ceil_div(n * ld, 256), 256, 0, s>>>(src, n, d, dst, ld);
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}

}  // namespace fir
