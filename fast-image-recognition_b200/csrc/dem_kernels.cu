// dem_kernels.cu — directed-enumeration ANN (qt_cpp/ann.cpp:270-507, the PIVOT build) on the GPU.
//
// Build  (ctor, ann.cpp:302-342): a sequential farthest-point pivot chain.  Step ii computes the exact distance
//         row P[ii][j] = feature_distance(db[j], db[pivot_ii]) for every gallery row (HBM-bound stream of the
//         gallery), adds it to fp64 running far-sums (the reference recomputes the same sums from scratch in the
//         same order, :315-321 — identical bits), takes the arg-max as the next pivot and the minimum other-class
//         distance for the false-accept threshold.  The chain lives entirely on the device: the next pivot index is
//         read by the next step's kernels from device memory, so there is no host round trip per step.
// Search (recognize, :416-507), batched over queries:
//         pivot distances → early exit → likelihood[ν] = Σ_i (d_i − P[i][ν])² (fp32, sequential in i) → candidates
//         in ascending (likelihood, index) order until one is under the threshold or the budget is spent.
//         The ordered walk is evaluated in rounds of geometrically growing size: a two-pass 16+16-bit radix select
//         finds the likelihood key that closes the round, the round's candidates are collected in index order,
//         their exact distances are evaluated in parallel, and a per-query reduction reproduces what the
//         sequential walk would have returned (first hit in order, else first minimum in order) plus its
//         distanceCalcCount.
#include "fir_common.cuh"
#include "handles.hpp"
#include "sharded.hpp"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <set>

struct fir_dem {
    fir_gallery* g = nullptr;
    int n_pivots = 0, chain_rows = 0;
    float threshold = 0.f;
    std::vector<int32_t> pivots;       // chain_rows entries; the first n_pivots are used by search
    std::vector<float> min_other;      // chain_rows entries
    float* P_raw = nullptr;            // [n_pivots][n] true pivot distances (device)
    float* P_search = nullptr;         // [n_pivots][n] with -1 where the reference's index walk skips (ν, step)
    unsigned char* in_tail = nullptr;  // [n] 1 = ν is a candidate after the pivot phase
    int32_t* d_pivots = nullptr;       // [n_pivots]
    int64_t tail_size = 0;
    // pivot-space gallery for the first round on tensor cores (see "round 0" below); absent for small galleries / S != 32
    fir_gallery* pt = nullptr;         // rows ν = (P[0][ν] .. P[S-1][ν]), L2 — its exact rerank IS the reference's likelihood sum
    float* center = nullptr;           // [32] per-pivot mean of P (shadow centring)
    unsigned char* exclude = nullptr;  // [n] rows outside the tensor round: the special rows below
    int n_special = 0;                 // rows touched by the reference's index walk (positions < S and the pivots themselves)
    int32_t* sp_rows = nullptr;        // [n_special]
    uint32_t* sp_mask = nullptr;       // [n_special] bit i = step i contributes to the row's likelihood
    unsigned char* sp_tail = nullptr;  // [n_special] 1 = the row is a candidate
    int32_t last_launches = 0;         // kernels launched by the last fir_dem_search
    // row-sharded build (fir_shard_dem_build): this handle covers gallery rows [row_lo, row_lo + g->n) of n_total; pivots,
    // QState indices and the candidate order are GLOBAL; the pivots' feature rows are replicated on every rank
    bool sharded = false;
    fir_comm* comm = nullptr;          // borrowed
    int64_t row_lo = 0, n_total = 0, min_shard_rows = 0;
    float* pivot_rows = nullptr;       // [n_pivots][dp]
};

namespace fir {

const fir_gallery* dem_gallery(const fir_dem* dem) { return dem ? dem->g : nullptr; }

// ---------------------------------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------------------------------
constexpr int RB = 256;   // threads per block of the reduction kernels

// far-sum update + block-level arg-max / min-other for one chain step
// (row shards: *cur_pivot is a GLOBAL row, `lo` the global index of local row 0, and the pivot's class comes from the
//  broadcast pivot record — the pivot itself may live on another rank)
__global__ void __launch_bounds__(RB) dem_step_kernel(const float* __restrict__ row, int64_t n, const int32_t* __restrict__ labels,
                                                      const int32_t* __restrict__ cur_pivot, double* __restrict__ far_sum,
                                                      float* __restrict__ keep_row, double* __restrict__ blk_max, int32_t* __restrict__ blk_arg,
                                                      float* __restrict__ blk_min, int64_t lo, const uint32_t* __restrict__ pivot_label) {
    __shared__ double s_max[RB];
    __shared__ int32_t s_arg[RB];
    __shared__ float s_min[RB];
    const int64_t pivot = (int64_t)*cur_pivot - lo;
    const int pl = pivot_label ? (int)*pivot_label : labels[pivot];
    double best = 0.0; int32_t arg = -1;                     // maxFarDist = 0, mostFarModel = -1  (ann.cpp:305-306)
    float mn = 3.402823466e+38f;                             // numeric_limits<float>::max()      (:307)
    for (int64_t j = (int64_t)blockIdx.x * RB + threadIdx.x; j < n; j += (int64_t)gridDim.x * RB) {
        const float d = row[j];
        if (keep_row) keep_row[j] = d;                       // :311
        if (labels[j] != pl && d < mn) mn = d;               // :312-314
        double s = (j == pivot) ? -1000000.0 : far_sum[j] + (double)d;   // :317-320
        far_sum[j] = s;
        if (s > best) { best = s; arg = (int32_t)j; }        // :322-325 (strict '>' ⇒ lowest j; j ascends within a thread)
    }
    s_max[threadIdx.x] = best; s_arg[threadIdx.x] = arg; s_min[threadIdx.x] = mn;
    __syncthreads();
    for (int o = RB / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            double ob = s_max[threadIdx.x + o]; int32_t oa = s_arg[threadIdx.x + o];
            double mb = s_max[threadIdx.x]; int32_t ma = s_arg[threadIdx.x];
            if (oa >= 0 && (ma < 0 || ob > mb || (ob == mb && oa < ma))) { s_max[threadIdx.x] = ob; s_arg[threadIdx.x] = oa; }
            s_min[threadIdx.x] = fminf(s_min[threadIdx.x], s_min[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { blk_max[blockIdx.x] = s_max[0]; blk_arg[blockIdx.x] = s_arg[0]; blk_min[blockIdx.x] = s_min[0]; }
}

__global__ void __launch_bounds__(RB) dem_step_final_kernel(const double* __restrict__ blk_max, const int32_t* __restrict__ blk_arg,
                                                            const float* __restrict__ blk_min, int nblk, int step, int chain_rows,
                                                            int32_t* __restrict__ pivots, float* __restrict__ min_other,
                                                            int32_t* __restrict__ cur_pivot, int64_t lo, unsigned long long* __restrict__ trip) {
    __shared__ double s_max[RB];
    __shared__ int32_t s_arg[RB];
    __shared__ float s_min[RB];
    double best = 0.0; int32_t arg = -1; float mn = 3.402823466e+38f;
    for (int b = threadIdx.x; b < nblk; b += RB) {
        double ob = blk_max[b]; int32_t oa = blk_arg[b];
        if (oa >= 0 && (arg < 0 || ob > best || (ob == best && oa < arg))) { best = ob; arg = oa; }
        mn = fminf(mn, blk_min[b]);
    }
    s_max[threadIdx.x] = best; s_arg[threadIdx.x] = arg; s_min[threadIdx.x] = mn;
    __syncthreads();
    for (int o = RB / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            double ob = s_max[threadIdx.x + o]; int32_t oa = s_arg[threadIdx.x + o];
            double mb = s_max[threadIdx.x]; int32_t ma = s_arg[threadIdx.x];
            if (oa >= 0 && (ma < 0 || ob > mb || (ob == mb && oa < ma))) { s_max[threadIdx.x] = ob; s_arg[threadIdx.x] = oa; }
            s_min[threadIdx.x] = fminf(s_min[threadIdx.x], s_min[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (trip) {                                                             // row shard: this rank's (far-sum maximum, global row, minimum) goes to the exchange
            trip[0] = (unsigned long long)__double_as_longlong(s_max[0]);
            trip[1] = ((unsigned long long)(uint32_t)(s_arg[0] < 0 ? -1 : (int32_t)(s_arg[0] + lo)) << 32) | __float_as_uint(s_min[0]);
            return;
        }
        min_other[step] = s_min[0];                                             // ann.cpp:327
        if (step + 1 < chain_rows) { pivots[step + 1] = s_arg[0]; *cur_pivot = s_arg[0] < 0 ? 0 : s_arg[0]; }   // :328-330
    }
}

// ---- row-sharded build: the pivot's feature row + class travel as raw bits (non-owners contribute zeros, all-reduce(max)
// over uint32 returns the owner's bits exactly), every rank's (maximum, row, minimum) triple is all-gathered and combined in
// the reference's scan order (strict '>' over ascending global rows ⇒ lowest row on ties) --------------------------------
__global__ void dem_pivot_record_kernel(const float* __restrict__ rows, int ld, const int32_t* __restrict__ labels, int64_t lo, int64_t n,
                                        const int32_t* __restrict__ cur_pivot, uint32_t* __restrict__ rec /* [ld + 1] */) {
    const int64_t p = (int64_t)*cur_pivot - lo;
    const bool mine = p >= 0 && p < n;
    for (int c = threadIdx.x; c <= ld; c += blockDim.x)
        rec[c] = !mine ? 0u : (c < ld ? __float_as_uint(rows[p * ld + c]) : (uint32_t)labels[p]);
}
__global__ void dem_pick_global_kernel(const unsigned long long* __restrict__ trips, int world, int step, int chain_rows, int32_t* __restrict__ pivots,
                                       float* __restrict__ min_other, int32_t* __restrict__ cur_pivot) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double best = 0.0; int32_t arg = -1; float mn = 3.402823466e+38f;
    for (int r = 0; r < world; ++r) {                                           // ranks hold ascending row ranges
        const double v = __longlong_as_double((long long)trips[2 * r]);
        const int32_t a = (int32_t)(uint32_t)(trips[2 * r + 1] >> 32);
        const float m = __uint_as_float((uint32_t)trips[2 * r + 1]);
        if (a >= 0 && (arg < 0 || v > best)) { best = v; arg = a; }
        mn = fminf(mn, m);
    }
    min_other[step] = mn;
    if (step + 1 < chain_rows) { pivots[step + 1] = arg; *cur_pivot = arg < 0 ? 0 : arg; }
}

// ---------------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------------
struct QState {          // per query, device
    float best_dist;     // bestDistance
    int32_t best_idx;    // bestIndex (local row)
    int32_t count;       // distanceCalcCount
    int32_t done;        // 1 = result final (hit under threshold, or budget spent)
    int32_t below;       // isFoundLessThreshold
    uint32_t last_key;   // (key, idx) of the last candidate position consumed so far
    int32_t last_idx;
    int32_t started;     // 0 until the first round consumed something
};

// pivot phase (ann.cpp:441-446 + CHECK_FOR_BEST_DIST :390-400)
__global__ void dem_pivot_phase_kernel(const float* __restrict__ pd, int64_t nq, int S, const int32_t* __restrict__ pivots, float threshold,
                                       int M, QState* __restrict__ st) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    QState s;
    s.best_dist = 3.402823466e+38f; s.best_idx = -1; s.count = 0; s.done = 0; s.below = 0; s.last_key = 0; s.last_idx = -1; s.started = 0;
    for (int i = 0; i < S; ++i) {
        const float d = pd[q * S + i];
        ++s.count;
        if (d < s.best_dist) {
            s.best_dist = d; s.best_idx = pivots[i];
            if (d < threshold) { s.below = 1; s.done = 1; break; }
        }
    }
    if (!s.done && s.count >= M) s.done = 1;                  // while (distanceCalcCount < imageCountToCheck) never runs (:472)
    st[q] = s;
}

// likelihood[q][ν] = Σ_i fl((d_i − P[i][ν])²) over the steps whose P entry is ≥ 0 (ann.cpp:453-461); +inf outside the tail.
// One thread owns one gallery row and QPT queries: every P element read from HBM is used QPT times (the pass is bound by
// that read — S·N·4 bytes per group of QPT queries — and by the N·4-byte likelihood row it writes per query).
template <int QPT>
__global__ void __launch_bounds__(256) dem_likelihood_kernel(const float* __restrict__ pd, const int32_t* __restrict__ qlist, int nqc, int S,
                                                             const float* __restrict__ P, int64_t n, const unsigned char* __restrict__ in_tail,
                                                             float* __restrict__ lik, float* __restrict__ gmin, int64_t ng) {
    __shared__ __align__(16) float dq[32][QPT];                // [pivot][query]: four queries per broadcast LDS.128
    const int q0 = blockIdx.y * QPT;
    for (int t = threadIdx.x; t < QPT * 32; t += 256) {
        const int qq = t >> 5, i = t & 31;
        dq[i][qq] = (q0 + qq < nqc && i < S) ? pd[(int64_t)qlist[q0 + qq] * S + i] : 0.f;
    }
    __syncthreads();
    const int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool valid = v < n;
    float acc[QPT];
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) acc[qq] = 0.f;
    if (valid) {
        for (int i = 0; i < S; ++i) {
            const float m = P[(int64_t)i * n + v];
            if (m >= 0.f) {                                                     // :456
#pragma unroll
                for (int qq = 0; qq < QPT; qq += 4) {
                    const float4 dv = *reinterpret_cast<const float4*>(&dq[i][qq]);
                    float t;
                    t = __fsub_rn(dv.x, m); acc[qq + 0] = __fadd_rn(acc[qq + 0], __fmul_rn(t, t));   // :457-458
                    t = __fsub_rn(dv.y, m); acc[qq + 1] = __fadd_rn(acc[qq + 1], __fmul_rn(t, t));
                    t = __fsub_rn(dv.z, m); acc[qq + 2] = __fadd_rn(acc[qq + 2], __fmul_rn(t, t));
                    t = __fsub_rn(dv.w, m); acc[qq + 3] = __fadd_rn(acc[qq + 3], __fmul_rn(t, t));
                }
            }
        }
    }
    const bool tail = valid && in_tail[v] != 0;
    const float inf = __int_as_float(0x7f800000);
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) {
        const float lv = tail ? acc[qq] : inf;
        if (valid && q0 + qq < nqc) lik[(int64_t)(q0 + qq) * n + v] = lv;
        if (gmin) {
            // minimum of the 32 rows of this warp: a bound for the first round's select (dem_fast_* below)
            float m = lv;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if ((threadIdx.x & 31) == 0 && q0 + qq < nqc && (v >> 5) < ng) gmin[(int64_t)(q0 + qq) * ng + (v >> 5)] = m;
        }
    }
}

// ---- fast first round ------------------------------------------------------------------------------
// The round wants the `want` smallest (likelihood, row) pairs.  The want-th smallest of the per-warp minima (found by the
// same radix select over an array 32x smaller) is an upper bound T for the want-th smallest likelihood: at least `want`
// rows are <= T, and rows <= T only live in warps whose minimum is <= T, so there are at most 32 x (want + ties) of them.
// One pass collects them, a shared-memory sort puts them in (likelihood, row) order and the first `want` ARE the round —
// one read of the likelihood rows instead of four.  Anything unusual (ties overflowing the buffer, fewer rows than wanted)
// leaves the query to the general path below.
constexpr int FAST_CAP = 8192 + 1024;

__global__ void __launch_bounds__(256) dem_fast_collect_kernel(const float* __restrict__ lik, const int32_t* __restrict__ qlist, int nqc, int64_t n,
                                                               const QState* __restrict__ st, const int32_t* __restrict__ t_all,
                                                               const uint32_t* __restrict__ t_key, uint32_t* __restrict__ fkey,
                                                               int32_t* __restrict__ frow, int32_t* __restrict__ fcnt) {
    const int qc = blockIdx.y;
    if (st[qlist[qc]].done || t_all[qc]) return;
    const float* lr = lik + (int64_t)qc * n;
    const uint32_t T = t_key[qc];
    for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < n; v += (int64_t)gridDim.x * 256) {
        const uint32_t key = __float_as_uint(lr[v]);
        if (key >= 0x7f800000u || key > T) continue;
        const int pos = atomicAdd(&fcnt[qc], 1);
        if (pos < FAST_CAP) { fkey[(int64_t)qc * FAST_CAP + pos] = key; frow[(int64_t)qc * FAST_CAP + pos] = (int32_t)v; }
    }
}

// one block per query: sort the collected (key, row) pairs, emit the first `want` as the round's candidates
__global__ void __launch_bounds__(256) dem_fast_finalize_kernel(const int32_t* __restrict__ qlist, int nqc, const QState* __restrict__ st,
                                                                const int32_t* __restrict__ t_all, const int32_t* __restrict__ want,
                                                                const uint32_t* __restrict__ fkey, const int32_t* __restrict__ frow,
                                                                const int32_t* __restrict__ fcnt, int cap, int32_t* __restrict__ cand,
                                                                uint32_t* __restrict__ cand_key, int32_t* __restrict__ round_n,
                                                                uint32_t* __restrict__ key_sel, int32_t* __restrict__ last_tie,
                                                                int32_t* __restrict__ take_all, int32_t* __restrict__ fast_ok) {
    extern __shared__ unsigned long long fsort[];
    const int qc = blockIdx.x;
    const int tid = threadIdx.x;
    const int c = fcnt[qc], w = want[qc];
    if (st[qlist[qc]].done || t_all[qc] || c > FAST_CAP || c < w || w <= 0 || w > cap) { if (tid == 0) fast_ok[qc] = 0; return; }
    int P2 = 1;
    while (P2 < c) P2 <<= 1;
    for (int i = tid; i < P2; i += 256)
        fsort[i] = i < c ? (((unsigned long long)fkey[(int64_t)qc * FAST_CAP + i] << 32) | (uint32_t)frow[(int64_t)qc * FAST_CAP + i]) : ~0ull;
    __syncthreads();
    for (int k2 = 2; k2 <= P2; k2 <<= 1)
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P2; i += 256) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = fsort[i], b = fsort[l];
                    const bool up = (i & k2) == 0;
                    if ((a > b) == up) { fsort[i] = b; fsort[l] = a; }
                }
            }
            __syncthreads();
        }
    int32_t* out = cand + (int64_t)qc * cap;
    uint32_t* outk = cand_key + (int64_t)qc * cap;
    for (int i = tid; i < cap; i += 256) {
        if (i < w) { out[i] = (int32_t)(uint32_t)fsort[i]; outk[i] = (uint32_t)(fsort[i] >> 32); }
        else out[i] = -1;
    }
    if (tid == 0) {
        round_n[qc] = w;
        key_sel[qc] = (uint32_t)(fsort[w - 1] >> 32);       // where the round ends in (key, row) order
        last_tie[qc] = (int32_t)(uint32_t)fsort[w - 1];
        take_all[qc] = 0;
        fast_ok[qc] = 1;
    }
}

__device__ __forceinline__ bool after_last(uint32_t key, int32_t idx, const QState& s) {
    return !s.started || key > s.last_key || (key == s.last_key && idx > s.last_idx);
}

// Radix select of the round's closing likelihood key, 11 + 11 + 10 bits.  `prefix` holds the digits chosen so far.
// Histograms are privatised in shared memory (the keys cluster in a few hundred bins — global atomics on 64 Ki bins spent
// 3 ms per pass on contention) and only the non-empty bins are flushed.
__device__ __forceinline__ bool radix_match(uint32_t key, int pass, uint32_t prefix) {
    return pass == 0 || (pass == 1 ? (key >> 21) == prefix : (key >> 10) == prefix);
}
__device__ __forceinline__ uint32_t radix_digit(uint32_t key, int pass) {
    return pass == 0 ? (key >> 21) : (pass == 1 ? ((key >> 10) & 0x7ffu) : (key & 0x3ffu));
}

__global__ void __launch_bounds__(256) dem_hist_kernel(const float* __restrict__ lik, const int32_t* __restrict__ qlist, int nqc, int64_t n,
                                                       const QState* __restrict__ st, int pass, const uint32_t* __restrict__ prefix,
                                                       const int32_t* __restrict__ take_all, uint32_t* __restrict__ hist, const int32_t* __restrict__ skip,
                                                       int32_t ioff = 0 /* row shards: QState rows are global, v is local */) {
    __shared__ uint32_t sh[2048];
    const int qc = blockIdx.y;
    if (skip && skip[qc]) return;
    const QState s = st[qlist[qc]];
    if (s.done || (pass > 0 && take_all[qc])) return;
    for (int i = threadIdx.x; i < 2048; i += 256) sh[i] = 0;
    __syncthreads();
    const float* lr = lik + (int64_t)qc * n;
    const uint32_t pre = pass ? prefix[qc] : 0;
    for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < n; v += (int64_t)gridDim.x * 256) {
        const uint32_t key = __float_as_uint(lr[v]);
        if (key >= 0x7f800000u) continue;                    // +inf: not in the tail
        if (!after_last(key, (int32_t)v + ioff, s)) continue;
        if (radix_match(key, pass, pre)) atomicAdd(&sh[radix_digit(key, pass)], 1u);
    }
    __syncthreads();
    uint32_t* h = hist + (size_t)qc * 2048;
    for (int i = threadIdx.x; i < 2048; i += 256)
        if (sh[i]) atomicAdd(&h[i], sh[i]);
}

// one block per query: walk the bins until the cumulative count reaches what is still wanted
__global__ void __launch_bounds__(32) dem_pick_kernel(const uint32_t* __restrict__ hist, const int32_t* __restrict__ qlist, int nqc,
                                                      const QState* __restrict__ st, int pass, const int32_t* __restrict__ want_in,
                                                      uint32_t* __restrict__ prefix, int32_t* __restrict__ want_rem, int32_t* __restrict__ take_all,
                                                      uint32_t* __restrict__ key_sel, int32_t* __restrict__ tie_take, int32_t* __restrict__ round_n,
                                                      const int32_t* __restrict__ skip) {
    const int qc = blockIdx.x;
    if (threadIdx.x != 0 || (skip && skip[qc])) return;
    const QState s = st[qlist[qc]];
    if (s.done) { round_n[qc] = 0; take_all[qc] = 0; key_sel[qc] = 0; tie_take[qc] = 0; return; }
    if (pass > 0 && take_all[qc]) return;
    const uint32_t* h = hist + (size_t)qc * 2048;
    const int want = pass == 0 ? want_in[qc] : want_rem[qc];
    const int bins = pass == 2 ? 1024 : 2048;
    uint32_t cum = 0; int bin = -1;
    for (int b = 0; b < bins; ++b) { if (cum + h[b] >= (uint32_t)want) { bin = b; break; } cum += h[b]; }
    if (pass == 0) {
        if (bin < 0) {                                       // fewer than `want` keys remain: the round takes everything that is left
            take_all[qc] = 1; key_sel[qc] = 0xffffffffu; tie_take[qc] = 0; round_n[qc] = (int32_t)cum;
            return;
        }
        take_all[qc] = 0;
        prefix[qc] = (uint32_t)bin; want_rem[qc] = want - (int32_t)cum;
    } else if (pass == 1) {
        prefix[qc] = (prefix[qc] << 11) | (uint32_t)bin; want_rem[qc] = want - (int32_t)cum;
    } else {
        key_sel[qc] = (prefix[qc] << 10) | (uint32_t)bin;
        tie_take[qc] = want - (int32_t)cum;                  // how many keys equal to key_sel close the round, lowest index first
        round_n[qc] = want_in[qc];
    }
}

// Parallel collection: everything below the closing key goes straight into the round's list (order is irrelevant, the
// reduction compares (key, index) explicitly); rows AT the closing key go to a small side buffer so that exactly the
// tie_take lowest indices can be taken afterwards.
constexpr int TIE_CAP = 32;
__global__ void __launch_bounds__(256) dem_collect_fast_kernel(const float* __restrict__ lik, const int32_t* __restrict__ qlist, int nqc, int64_t n,
                                                               const QState* __restrict__ st, const int32_t* __restrict__ take_all,
                                                               const uint32_t* __restrict__ key_sel, int cap, int32_t* __restrict__ cand,
                                                               uint32_t* __restrict__ cand_key, int32_t* __restrict__ cnt,
                                                               int32_t* __restrict__ tie_buf, int32_t* __restrict__ tie_cnt, const int32_t* __restrict__ skip,
                                                               int32_t ioff = 0) {
    const int qc = blockIdx.y;
    if (skip && skip[qc]) return;
    const QState s = st[qlist[qc]];
    if (s.done) return;
    const float* lr = lik + (int64_t)qc * n;
    const bool all = take_all[qc] != 0;
    const uint32_t ksel = key_sel[qc];
    int32_t* out = cand + (int64_t)qc * cap;
    uint32_t* outk = cand_key + (int64_t)qc * cap;
    for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < n; v += (int64_t)gridDim.x * 256) {
        const uint32_t key = __float_as_uint(lr[v]);
        if (key >= 0x7f800000u || !after_last(key, (int32_t)v + ioff, s)) continue;
        if (all || key < ksel) {
            const int pos = atomicAdd(&cnt[qc], 1);
            if (pos < cap) { out[pos] = (int32_t)v; outk[pos] = key; }
        } else if (key == ksel) {
            const int pos = atomicAdd(&tie_cnt[qc], 1);
            if (pos < TIE_CAP) tie_buf[qc * TIE_CAP + pos] = (int32_t)v;
        }
    }
}

// one thread per query: append the tie_take lowest-index ties, pad the list; flags queries whose ties overflowed the buffer
__global__ void dem_collect_fixup_kernel(const int32_t* __restrict__ qlist, int nqc, const QState* __restrict__ st, const int32_t* __restrict__ take_all,
                                         const uint32_t* __restrict__ key_sel, const int32_t* __restrict__ tie_take, int cap,
                                         int32_t* __restrict__ cand, uint32_t* __restrict__ cand_key, int32_t* __restrict__ cnt,
                                         int32_t* __restrict__ tie_buf, const int32_t* __restrict__ tie_cnt, int32_t* __restrict__ last_tie_idx,
                                         int32_t* __restrict__ overflow, const int32_t* __restrict__ skip) {
    const int qc = blockIdx.x * blockDim.x + threadIdx.x;
    if (qc >= nqc) return;
    overflow[qc] = 0;
    if (st[qlist[qc]].done || (skip && skip[qc])) return;
    int32_t* out = cand + (int64_t)qc * cap;
    uint32_t* outk = cand_key + (int64_t)qc * cap;
    int c = min(cnt[qc], cap);
    int ltie = -1;
    if (!take_all[qc]) {
        const int nt = tie_cnt[qc];
        if (nt > TIE_CAP) { overflow[qc] = 1; return; }       // rare: mass ties at the closing key → ordered slow path
        int32_t* tb = tie_buf + qc * TIE_CAP;
        for (int i = 1; i < nt; ++i) {                        // insertion sort by index
            const int32_t x = tb[i]; int j = i - 1;
            while (j >= 0 && tb[j] > x) { tb[j + 1] = tb[j]; --j; }
            tb[j + 1] = x;
        }
        const int take = min(tie_take[qc], nt);
        for (int i = 0; i < take && c < cap; ++i) { out[c] = tb[i]; outk[c] = key_sel[qc]; ltie = tb[i]; ++c; }
    }
    for (int i = c; i < cap; ++i) out[i] = -1;
    last_tie_idx[qc] = ltie;
}

// one warp per query: collect the round's candidates in INDEX order (ballot compaction keeps the order)
__global__ void __launch_bounds__(128) dem_collect_kernel(const float* __restrict__ lik, const int32_t* __restrict__ qlist, int nqc, int64_t n,
                                                          const QState* __restrict__ st, const int32_t* __restrict__ overflow,
                                                          const uint32_t* __restrict__ key_sel, const int32_t* __restrict__ tie_take,
                                                          int cap, int32_t* __restrict__ cand, uint32_t* __restrict__ cand_key,
                                                          int32_t* __restrict__ last_tie_idx, int32_t ioff = 0) {
    const int lane = threadIdx.x & 31;
    const int qc = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (qc >= nqc) return;
    const QState s = st[qlist[qc]];
    int32_t* out = cand + (int64_t)qc * cap;
    uint32_t* outk = cand_key + (int64_t)qc * cap;
    if (s.done || !overflow[qc]) return;                     // slow path: only queries whose closing-key ties overflowed the side buffer
    const float* lr = lik + (int64_t)qc * n;
    const bool take_all = key_sel[qc] == 0xffffffffu;
    const uint32_t ksel = key_sel[qc];
    int ties_left = tie_take[qc];
    int cnt = 0, ltie = -1;
    for (int64_t base = 0; base < n; base += 32) {
        const int64_t v = base + lane;
        bool ok = false, tie = false;
        uint32_t key = 0;
        if (v < n) {
            key = __float_as_uint(lr[v]);
            if (key < 0x7f800000u && after_last(key, (int32_t)v + ioff, s)) {
                if (take_all || key < ksel) ok = true;
                else if (key == ksel) tie = true;
            }
        }
        const uint32_t tmask = __ballot_sync(0xffffffffu, tie);
        if (tie) {                                                              // the first `ties_left` ties in index order are in the round
            const int rank = __popc(tmask & ((1u << lane) - 1));
            if (rank < ties_left) ok = true;
        }
        const int taken_ties = min(__popc(tmask), ties_left);
        if (taken_ties > 0) {
            // index of the last tie taken in this group
            uint32_t mm = tmask; int seen = 0, li = -1;
            while (mm && seen < taken_ties) { li = __ffs(mm) - 1; mm &= mm - 1; ++seen; }
            ltie = (int)(base + li);
        }
        ties_left -= taken_ties;
        const uint32_t omask = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const int pos = cnt + __popc(omask & ((1u << lane) - 1));
            if (pos < cap) { out[pos] = (int32_t)v; outk[pos] = key; }
        }
        cnt += __popc(omask);
    }
    for (int i = cnt + lane; i < cap; i += 32) out[i] = -1;
    if (lane == 0) last_tie_idx[qc] = ltie;
}

// one warp per query: fold the round's exact distances into the query state exactly as the sequential walk would
__global__ void __launch_bounds__(128) dem_reduce_kernel(const float* __restrict__ cdist, const int32_t* __restrict__ cand,
                                                         const uint32_t* __restrict__ cand_key, const int32_t* __restrict__ qlist, int nqc,
                                                         int cap, const int32_t* __restrict__ round_n, const uint32_t* __restrict__ key_sel,
                                                         const int32_t* __restrict__ last_tie_idx, float threshold, int M, int64_t tail_size, int S,
                                                         QState* __restrict__ st, int32_t* __restrict__ n_active) {
    const int lane = threadIdx.x & 31;
    const int qc = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (qc >= nqc) return;
    const int qi = qlist ? qlist[qc] : qc;
    QState s = st[qi];
    if (s.done || round_n[qc] < 0) return;
    const int m = min(round_n[qc], cap);
    const int32_t* ci = cand + (int64_t)qc * cap;
    const uint32_t* ck = cand_key + (int64_t)qc * cap;
    const float* cd = cdist + (int64_t)qc * cap;
    // first hit in (key, idx) order; otherwise the first minimum in (dist, key, idx) order
    uint32_t hk = 0xffffffffu; int32_t hi = 0x7fffffff; float hd = 0.f; bool hit = false;
    float bd = 3.402823466e+38f; uint32_t bk = 0xffffffffu; int32_t bi = 0x7fffffff; bool any = false;
    for (int i = lane; i < m; i += 32) {
        const float d = cd[i]; const uint32_t k = ck[i]; const int32_t v = ci[i];
        if (d < threshold && (!hit || k < hk || (k == hk && v < hi))) { hit = true; hk = k; hi = v; hd = d; }
        if (!any || d < bd || (d == bd && (k < bk || (k == bk && v < bi)))) { any = true; bd = d; bk = k; bi = v; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const int oh = __shfl_xor_sync(0xffffffffu, (int)hit, o);
        const uint32_t ohk = __shfl_xor_sync(0xffffffffu, hk, o); const int32_t ohi = __shfl_xor_sync(0xffffffffu, hi, o);
        const float ohd = __shfl_xor_sync(0xffffffffu, hd, o);
        if (oh && (!hit || ohk < hk || (ohk == hk && ohi < hi))) { hit = true; hk = ohk; hi = ohi; hd = ohd; }
        const int oa = __shfl_xor_sync(0xffffffffu, (int)any, o);
        const float obd = __shfl_xor_sync(0xffffffffu, bd, o); const uint32_t obk = __shfl_xor_sync(0xffffffffu, bk, o);
        const int32_t obi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oa && (!any || obd < bd || (obd == bd && (obk < bk || (obk == bk && obi < bi))))) { any = true; bd = obd; bk = obk; bi = obi; }
    }
    if (hit) {
        // While walking, bestDistance ≥ threshold (else the walk would already have stopped), so the first candidate in
        // order that is under the threshold also beats bestDistance: it is the answer (CHECK_FOR_BEST_DIST, ann.cpp:390-400).
        int rank = 0;
        for (int i = lane; i < m; i += 32) {
            const uint32_t k = ck[i]; const int32_t v = ci[i];
            if (k < hk || (k == hk && v < hi)) ++rank;
        }
        for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
        if (lane == 0) {
            s.best_dist = hd; s.best_idx = hi; s.below = 1; s.done = 1;
            s.count += rank + 1;
            st[qi] = s;
        }
        return;
    }
    if (lane == 0) {
        if (any && bd < s.best_dist) { s.best_dist = bd; s.best_idx = bi; }
        s.count += m;
        s.started = 1;
        // position of the end of this round in (key, idx) order
        if (key_sel[qc] == 0xffffffffu) { s.done = 1; }                          // tail exhausted
        else {
            s.last_key = key_sel[qc];
            s.last_idx = last_tie_idx[qc];
        }
        if (s.count >= M || (int64_t)(s.count - S) >= tail_size) s.done = 1;
        if (!s.done) atomicAdd(n_active, 1);
        st[qi] = s;
    }
}

// ---- row shards: the round's exchange ---------------------------------------------------------------
// Every rank walks ITS rows: the local next-`want` candidates after the query's (last_key, last_idx) in (likelihood, row)
// order, with their exact distances, packed as records {likelihood bits, GLOBAL row, distance bits}.  The records of all
// ranks are all-gathered; the global first `want` of the union are the first `want` of the whole gallery's walk (each rank
// supplied its own first `want`), so sorting the union and folding it in order reproduces recognize() (ann.cpp:472-476) —
// identically on every rank, which keeps the replicated query states in step.
__global__ void dem_pack_records_kernel(const int32_t* __restrict__ cand, const uint32_t* __restrict__ cand_key, const float* __restrict__ cdist,
                                        const int32_t* __restrict__ round_n, int nqc, int cap, int32_t ioff, uint32_t* __restrict__ rec) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nqc * cap) return;
    const int qc = (int)(t / cap), j = (int)(t - (int64_t)qc * cap);
    const int32_t v = cand[t];
    const bool live = v >= 0 && j < max(round_n[qc], 0);
    rec[3 * t + 0] = live ? cand_key[t] : 0xffffffffu;
    rec[3 * t + 1] = live ? (uint32_t)(v + ioff) : 0xffffffffu;
    rec[3 * t + 2] = live ? __float_as_uint(cdist[t]) : 0u;
}

// one block per query: bitonic sort of the union by (likelihood, global row), then the in-order fold
__global__ void __launch_bounds__(256) dem_global_reduce_kernel(const uint32_t* __restrict__ rec /* [world][nqc][cap][3] */, int world, int nqc, int cap,
                                                                const int32_t* __restrict__ qlist, int round_size, float threshold, int M,
                                                                int64_t tail_size, int S, QState* __restrict__ st, int32_t* __restrict__ n_active) {
    extern __shared__ unsigned long long gsort[];          // [P2] keys, then [P2] uint32 distance bits
    __shared__ int s_hit_pos;
    __shared__ unsigned long long s_best;                  // (ordered distance bits << 32) | position: first minimum in order
    const int qc = blockIdx.x, tid = threadIdx.x;
    const int q = qlist ? qlist[qc] : qc;
    QState s = st[q];
    if (s.done) return;
    const int total = world * cap;
    int P2 = 1;
    while (P2 < total) P2 <<= 1;
    uint32_t* gdist = reinterpret_cast<uint32_t*>(gsort + P2);
    // sort key = (likelihood bits << 32 | global row); the distance rides along through an index sort: keys are unique per
    // live record (a row appears once), so the payload is looked up again by binary position after sorting pairs together
    for (int i = tid; i < P2; i += 256) {
        unsigned long long k = ~0ull; uint32_t d = 0u;
        if (i < total) {
            const int r = i / cap, j = i - r * cap;
            const uint32_t* e = rec + 3 * (((int64_t)r * nqc + qc) * cap + j);
            if (e[1] != 0xffffffffu) { k = ((unsigned long long)e[0] << 32) | e[1]; d = e[2]; }
        }
        gsort[i] = k; gdist[i] = d;
    }
    if (tid == 0) { s_hit_pos = 0x7fffffff; s_best = ~0ull; }
    __syncthreads();
    for (int k2 = 2; k2 <= P2; k2 <<= 1)
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P2; i += 256) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = gsort[i], b = gsort[l];
                    const bool up = (i & k2) == 0;
                    if ((a > b) == up) { gsort[i] = b; gsort[l] = a; const uint32_t t = gdist[i]; gdist[i] = gdist[l]; gdist[l] = t; }
                }
            }
            __syncthreads();
        }
    const int want = max(0, min(round_size, M - s.count));
    // live records are a prefix after sorting; count them (binary search by thread 0 is enough: P2 <= 16384)
    int valid;
    { int lo = 0, hi = total; while (lo < hi) { const int mid = (lo + hi) >> 1; if (gsort[mid] != ~0ull) lo = mid + 1; else hi = mid; } valid = lo; }
    const int m = min(want, valid);
    for (int i = tid; i < m; i += 256) {
        const float d = __uint_as_float(gdist[i]);
        if (d < threshold) atomicMin(&s_hit_pos, i);
        atomicMin(&s_best, ((unsigned long long)ordered_bits(d) << 32) | (uint32_t)i);
    }
    __syncthreads();
    if (tid != 0) return;
    if (s_hit_pos != 0x7fffffff) {                          // CHECK_FOR_BEST_DIST (ann.cpp:390-400): the first candidate under the threshold ends the walk
        const int p = s_hit_pos;
        s.best_dist = __uint_as_float(gdist[p]); s.best_idx = (int32_t)(uint32_t)gsort[p]; s.below = 1; s.done = 1;
        s.count += p + 1;
        st[q] = s;
        return;
    }
    if (m > 0) {
        const int p = (int)(uint32_t)s_best;
        const float bd = __uint_as_float(gdist[p]);
        if (bd < s.best_dist) { s.best_dist = bd; s.best_idx = (int32_t)(uint32_t)gsort[p]; }
        s.count += m;
        s.started = 1;
        s.last_key = (uint32_t)(gsort[m - 1] >> 32);
        s.last_idx = (int32_t)(uint32_t)gsort[m - 1];
    }
    if (valid < want) s.done = 1;                           // every rank handed over all it had left: the tail is exhausted
    if (s.count >= M || (int64_t)(s.count - S) >= tail_size) s.done = 1;
    if (!s.done) atomicAdd(n_active, 1);
    st[q] = s;
}

__global__ void dem_fill_pivot_cand_kernel(const int32_t* __restrict__ pivots, int S, int64_t nq, int32_t* __restrict__ cand) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq * S) cand[i] = pivots[i % S];
}
__global__ void dem_want_kernel(const QState* __restrict__ st, const int32_t* __restrict__ qlist, int nqc, int round_size, int M, int32_t* want) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nqc) return;
    const QState s = st[qlist[i]];
    int w = s.done ? 0 : min(round_size, M - s.count);
    want[i] = max(w, 0);
}
__global__ void dem_output_kernel(const QState* __restrict__ st, int64_t nq, int64_t index_offset, int32_t* idx, float* dist, uint8_t* below, int32_t* evals) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const QState s = st[q];
    idx[q] = s.best_idx < 0 ? -1 : (int32_t)(s.best_idx + index_offset);
    if (dist) dist[q] = s.best_dist;
    if (below) below[q] = (uint8_t)s.below;
    if (evals) evals[q] = s.count;
}
// The unfinished queries in ASCENDING order (one block walks the queries 1024 at a time; ballot + prefix counts keep the order).
// The order matters for row shards: position qc of the list names the same query on every rank, so the records exchanged
// for qc belong together (an atomic-append list would be permuted differently on each GPU).
__global__ void __launch_bounds__(1024) dem_active_list_kernel(const QState* __restrict__ st, int64_t nq, int32_t* list, int32_t* count) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int64_t q0 = 0; q0 < nq; q0 += 1024) {
        const int64_t q = q0 + threadIdx.x;
        const bool act = q < nq && !st[q].done;
        const uint32_t m = __ballot_sync(0xffffffffu, act);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (act) list[s_base + before + __popc(m & ((1u << lane) - 1))] = (int32_t)q;
        __syncthreads();
        if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < 32; ++w) tot += s_warp[w]; s_base += tot; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = s_base;
}

// ---- round 0 on tensor cores -----------------------------------------------------------------------
// likelihood[ν] = Σ_i (d_i − P[i][ν])² is the squared Euclidean distance, in the 32-dimensional "pivot space", between the
// query's pivot-distance vector (d_0..d_31) and column ν of P: the candidate order of recognize() (ann.cpp:453-476) is a
// nearest-neighbour order.  So the first round is ONE call of the tensor-core brute force (l2_tensor.cu) over the derived
// gallery Pᵀ [N][32]: tcgen05 candidates (q·xᵀ over K = 32), exact fp32 rerank — whose sequential sum fl(acc + fl(t·t)),
// t = fl(d_i − P_iν), is bit for bit the reference's likelihood accumulation (:457-458; the mean's division by 32 is exact)
// — certificate, top-k by (likelihood, row).  The [Q][N] likelihood matrix is never formed: per (query, row) the work is
// one accumulator element of a K = 32 MMA instead of 96 rounded FP32 instructions.  The ≤ 64 rows whose likelihood the
// reference's index walk truncates (pivots and positions < S) are excluded from the derived gallery and merged in exactly
// from a side list.  Queries that need more than the first round continue in the general path below from the round's
// closing (likelihood, row) key.
constexpr int kRound0K = 24;            // candidates of the first round (the tensor path serves k <= 28)
constexpr int64_t kDemTensorMinRows = 8192;

__global__ void dem_center_kernel(const float* __restrict__ P, int64_t n, float* __restrict__ center) {
    __shared__ double sh[256];
    const float* row = P + (int64_t)blockIdx.x * n;
    double a = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += 256) a += (double)row[j];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) center[blockIdx.x] = (float)(sh[0] / (double)n);
}
// PT[ν][i] = P[i][ν]; excluded rows get a far sentinel so that even the exact CUDA-core fallback never lists them (≥ 100000)
__global__ void dem_transpose_kernel(const float* __restrict__ P, int64_t n, int S, const unsigned char* __restrict__ exclude, float* __restrict__ PT) {
    __shared__ float tile[32][33];
    const int64_t v0 = (int64_t)blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8) { const int64_t v = v0 + threadIdx.x; tile[i][threadIdx.x] = (i < S && v < n) ? P[(int64_t)i * n + v] : 0.f; }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int64_t v = v0 + r;
        if (v < n) PT[v * 32 + threadIdx.x] = exclude[v] ? (threadIdx.x < S ? 1e18f : 0.f) : tile[threadIdx.x][r];
    }
}

// one warp per query: the round's candidates = the first `want` of (tensor top-k over the ordinary rows) ∪ (special rows with
// their truncated likelihoods, computed here), in (likelihood, row) order
__global__ void __launch_bounds__(128) dem_round0_merge_kernel(const int32_t* __restrict__ t_idx, const float* __restrict__ t_dist, int k0, int64_t nq,
                                                               const float* __restrict__ pd, int S, const float* __restrict__ P, int64_t n,
                                                               const int32_t* __restrict__ sp_rows, const uint32_t* __restrict__ sp_mask,
                                                               const unsigned char* __restrict__ sp_tail, int n_special, int M,
                                                               const QState* __restrict__ st, int32_t* __restrict__ cand, uint32_t* __restrict__ cand_key,
                                                               int32_t* __restrict__ round_n, uint32_t* __restrict__ key_sel, int32_t* __restrict__ last_tie) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q >= nq) return;
    const QState s = st[q];
    int32_t* out = cand + q * k0;
    uint32_t* outk = cand_key + q * k0;
    if (s.done) {
        if (lane < k0) out[lane] = -1;
        if (lane == 0) { round_n[q] = 0; key_sel[q] = 0; last_tie[q] = -1; }
        return;
    }
    const int want = min(k0, M - s.count);
    unsigned long long e0 = ~0ull, e1 = ~0ull, e2 = ~0ull;
    bool tiny = false;
    if (lane < k0) {
        const int32_t idx = t_idx[q * k0 + lane];
        if (idx >= 0) {
            const float dd = t_dist[q * k0 + lane];
            tiny = dd != 0.f && dd < 7.9e-31f;                   // fl(acc/32) may have lost bits below 2^-100: leave the query to the general path
            e0 = ((unsigned long long)__float_as_uint(__fmul_rn(dd, 32.f)) << 32) | (uint32_t)idx;
        }
    }
    for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
        if (j < n_special && sp_tail[j]) {
            const int32_t row = sp_rows[j];
            const uint32_t mask = sp_mask[j];
            float acc = 0.f;
            for (int i = 0; i < S; ++i)
                if ((mask >> i) & 1u) { const float t = __fsub_rn(pd[q * S + i], P[(int64_t)i * n + row]); acc = __fadd_rn(acc, __fmul_rn(t, t)); }
            const unsigned long long key = ((unsigned long long)__float_as_uint(acc) << 32) | (uint32_t)row;
            if (h == 0) e1 = key; else e2 = key;
        }
    }
    if (__any_sync(0xffffffffu, tiny)) {
        if (lane < k0) out[lane] = -1;
        if (lane == 0) { round_n[q] = -1; key_sel[q] = 0; last_tie[q] = -1; }
        return;
    }
    int taken = 0;
    unsigned long long last = 0;
    for (int r = 0; r < want; ++r) {
        unsigned long long mine = e0 < e1 ? e0 : e1;
        mine = mine < e2 ? mine : e2;
        unsigned long long best = mine;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long ov = __shfl_xor_sync(0xffffffffu, best, o); best = ov < best ? ov : best; }
        if (best == ~0ull) break;                                // fewer rows than wanted: the tail is exhausted
        if (mine == best) { if (e0 == best) e0 = ~0ull; else if (e1 == best) e1 = ~0ull; else e2 = ~0ull; }   // keys are distinct (one row each)
        if (lane == 0) { out[r] = (int32_t)(uint32_t)best; outk[r] = (uint32_t)(best >> 32); }
        last = best; ++taken;
    }
    for (int r = taken + lane; r < k0; r += 32) out[r] = -1;
    if (lane == 0) {
        round_n[q] = taken;
        key_sel[q] = taken < want ? 0xffffffffu : (uint32_t)(last >> 32);
        last_tie[q] = taken > 0 ? (int32_t)(uint32_t)last : -1;
    }
}

}  // namespace fir

using namespace fir;

// Replay the reference's likelihood_indices walk (ann.cpp:432-446) once on the host: it is query independent.
// li starts as the identity; step i does li[p_i] = li[i]; li[i] = p_i.  ν receives the likelihood term of step i iff it
// sits at a position > i afterwards; it is a candidate iff it sits at a position ≥ S at the end.
static int finalize_search_state(fir_dem* dm) {
    fir_gallery* g = dm->g;
    cudaStream_t s = g->stream;
    const int64_t n = g->n;
    const int S = dm->n_pivots;
    std::map<int64_t, int64_t> li;                       // sparse overrides of the identity
    auto get = [&](int64_t pos) { auto it = li.find(pos); return it == li.end() ? pos : it->second; };
    std::set<int64_t> special;
    for (int i = 0; i < S; ++i) { special.insert(i); special.insert(dm->pivots[i]); }
    std::vector<int64_t> sp(special.begin(), special.end());
    std::vector<std::vector<unsigned char> > member(S, std::vector<unsigned char>(sp.size(), 0));
    for (int i = 0; i < S; ++i) {
        const int64_t p = dm->pivots[i];
        li[p] = get(i);
        li[i] = p;
        for (size_t k = 0; k < sp.size(); ++k) {
            const int64_t v = sp[k];
            int mult = 0;
            for (auto& kv : li) if (kv.first > i && kv.second == v) ++mult;
            if (li.find(v) == li.end() && v > i) ++mult;                    // still at its identity position
            if (mult > 1) return fail(FIR_ERR_UNSUPPORTED, "pivot list repeats a pivot: the reference's index walk duplicates a candidate");
            member[i][k] = (unsigned char)mult;
        }
    }
    // (row shards: the walk above is over GLOBAL rows; this handle applies it to the special rows it holds and drops the rest)
    if (dm->sharded) {
        std::vector<int64_t> sp_local;
        std::vector<std::vector<unsigned char> > member_local(S);
        for (size_t k = 0; k < sp.size(); ++k)
            if (sp[k] >= dm->row_lo && sp[k] < dm->row_lo + n) {
                sp_local.push_back(sp[k] - dm->row_lo);
                for (int i = 0; i < S; ++i) member_local[i].push_back(member[i][k]);
            }
        sp.swap(sp_local);
        member.swap(member_local);
    }
    FIR_CUDA_TRY(cudaMemcpyAsync(dm->P_search, dm->P_raw, 4 * (size_t)S * n, cudaMemcpyDeviceToDevice, s));
    FIR_CUDA_TRY(cudaMemsetAsync(dm->in_tail, 1, (size_t)n, s));
    const float neg = -1.f; const unsigned char zero = 0;
    for (size_t k = 0; k < sp.size(); ++k) {
        for (int i = 0; i < S; ++i)
            if (!member[i][k]) FIR_CUDA_TRY(cudaMemcpyAsync(dm->P_search + (size_t)i * n + sp[k], &neg, 4, cudaMemcpyHostToDevice, s));
        if (!member[S - 1][k]) FIR_CUDA_TRY(cudaMemcpyAsync(dm->in_tail + sp[k], &zero, 1, cudaMemcpyHostToDevice, s));
    }
    dm->tail_size = (dm->sharded ? dm->n_total : n) - S; // each step moves exactly one position into the prefix
    FIR_CUDA_TRY(cudaMemcpyAsync(dm->d_pivots, dm->pivots.data(), 4 * (size_t)S, cudaMemcpyHostToDevice, s));
    FIR_CUDA_TRY(cudaStreamSynchronize(s));
    // round 0 on tensor cores: the pivot-space gallery (see dem_round0_merge_kernel)
    static const int tensor_on = [] { const char* e = getenv("FIR_DEM_TENSOR"); return e ? atoi(e) : 1; }();
    // (row shards: every rank must take the same decision — the rounds are collective — so it rests on the smallest shard)
    if (tensor_on && S == 32 && (dm->sharded ? dm->min_shard_rows : n) >= kDemTensorMinRows && g->cc_major >= 10 && tensor_path_supported(32)) {
        const int ns = (int)sp.size();
        std::vector<int32_t> rows(ns);
        std::vector<uint32_t> mask(ns, 0u);
        std::vector<unsigned char> tail(ns, 0), excl((size_t)n, 0);
        for (int k = 0; k < ns; ++k) {
            rows[k] = (int32_t)sp[k];
            for (int i = 0; i < S; ++i) if (member[i][k]) mask[k] |= 1u << i;
            tail[k] = member[S - 1][k];
            excl[(size_t)sp[k]] = 1;
        }
        dm->n_special = ns;
        float* PT = nullptr;
        auto drop = [&](int code) { cudaFree(PT); return code; };
        if (cudaMalloc(&dm->sp_rows, 4 * (size_t)ns) != cudaSuccess || cudaMalloc(&dm->sp_mask, 4 * (size_t)ns) != cudaSuccess ||
            cudaMalloc(&dm->sp_tail, (size_t)ns) != cudaSuccess || cudaMalloc(&dm->exclude, (size_t)n) != cudaSuccess ||
            cudaMalloc(&dm->center, 4 * 32) != cudaSuccess || cudaMalloc(&PT, 4 * (size_t)n * 32) != cudaSuccess)
            return drop(fail(FIR_ERR_OOM, "DEM pivot-space gallery allocation failed"));
        FIR_CUDA_TRY(cudaMemcpyAsync(dm->sp_rows, rows.data(), 4 * (size_t)ns, cudaMemcpyHostToDevice, s));
        FIR_CUDA_TRY(cudaMemcpyAsync(dm->sp_mask, mask.data(), 4 * (size_t)ns, cudaMemcpyHostToDevice, s));
        FIR_CUDA_TRY(cudaMemcpyAsync(dm->sp_tail, tail.data(), (size_t)ns, cudaMemcpyHostToDevice, s));
        FIR_CUDA_TRY(cudaMemcpyAsync(dm->exclude, excl.data(), (size_t)n, cudaMemcpyHostToDevice, s));
        dem_center_kernel<<<32, 256, 0, s>>>(dm->P_raw, n, dm->center);
        dem_transpose_kernel<<<(unsigned)ceil_div(n, 32), dim3(32, 8), 0, s>>>(dm->P_raw, n, S, dm->exclude, PT);
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) return drop(fail(FIR_ERR_CUDA, "DEM pivot-space gallery kernels failed"));
        const int st_ = fir_gallery_create(PT, nullptr, n, 32, FIR_L2, FIR_DEVICE, 0, &dm->pt);
        cudaFree(PT); PT = nullptr;
        if (st_ != FIR_OK) return st_;
        fir_gallery_set_stream(dm->pt, s);
        dm->pt->tensor_center = dm->center;
        dm->pt->tensor_exclude = dm->exclude;
    }
    return FIR_OK;
}

extern "C" {

int fir_dem_build(fir_gallery* g, const fir_dem_params* params, fir_dem** out) {
    if (!out) return fail(FIR_ERR_BAD_ARG, "out is null");
    *out = nullptr;
    if (!g || !params) return fail(FIR_ERR_BAD_ARG, "null argument");
    FIR_CUDA_TRY(cudaSetDevice(g->device));
    const int64_t n = g->n;
    if (n < 2) return fail(FIR_ERR_BAD_ARG, "directed enumeration needs at least 2 gallery rows");
    int np = (int)((double)n * 0.015);                              // ann.cpp:373  (int)(dbSize*0.015)
    if (np < 5) np = 5;                                             // :374-375
    if (params->max_chain > 0 && params->max_chain < np) np = params->max_chain;
    if (np > n) np = (int)n;
    int keep = params->max_pivots > 0 ? std::min(params->max_pivots, 32) : 32;   // :332-333 (search mask is 32 bits wide)
    keep = std::min(keep, np);
    int pivot0 = params->pivot0;
    if (pivot0 < 0) {                                               // stand-in for random_shuffle's head (:366-369)
        uint64_t z = (uint64_t)params->seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
        z ^= z >> 31; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 29;
        pivot0 = (int)(z % (uint64_t)n);
    }
    if (pivot0 >= n) return fail(FIR_ERR_BAD_ARG, "pivot0 out of range");

    fir_dem* dm = new fir_dem();
    dm->g = g;
    dm->chain_rows = np;
    dm->n_pivots = keep;
    cudaStream_t s = g->stream;
    const int nblk = (int)std::min<int64_t>(ceil_div(n, RB), 1184);
    float* row = nullptr; double* far_sum = nullptr; double* blk_max = nullptr; int32_t* blk_arg = nullptr; float* blk_min = nullptr;
    int32_t* d_chain = nullptr; float* d_min_other = nullptr; int32_t* cur = nullptr;
    auto cleanup = [&](int code) {
        cudaFree(row); cudaFree(far_sum); cudaFree(blk_max); cudaFree(blk_arg); cudaFree(blk_min); cudaFree(d_chain); cudaFree(d_min_other); cudaFree(cur);
        if (code != FIR_OK) fir_dem_destroy(dm);
        return code;
    };
    if (cudaMalloc(&row, 4 * (size_t)n) != cudaSuccess || cudaMalloc(&far_sum, 8 * (size_t)n) != cudaSuccess ||
        cudaMalloc(&blk_max, 8 * (size_t)nblk) != cudaSuccess || cudaMalloc(&blk_arg, 4 * (size_t)nblk) != cudaSuccess ||
        cudaMalloc(&blk_min, 4 * (size_t)nblk) != cudaSuccess || cudaMalloc(&d_chain, 4 * (size_t)np) != cudaSuccess ||
        cudaMalloc(&d_min_other, 4 * (size_t)np) != cudaSuccess || cudaMalloc(&cur, 4) != cudaSuccess ||
        cudaMalloc(&dm->P_raw, 4 * (size_t)keep * n) != cudaSuccess || cudaMalloc(&dm->P_search, 4 * (size_t)keep * n) != cudaSuccess ||
        cudaMalloc(&dm->in_tail, (size_t)n) != cudaSuccess || cudaMalloc(&dm->d_pivots, 4 * (size_t)keep) != cudaSuccess)
        return cleanup(fail(FIR_ERR_OOM, "DEM build allocation failed"));
    cudaMemsetAsync(far_sum, 0, 8 * (size_t)n, s);
    cudaMemsetAsync(d_chain, 0xFF, 4 * (size_t)np, s);
    cudaMemcpyAsync(d_chain, &pivot0, 4, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(cur, &pivot0, 4, cudaMemcpyHostToDevice, s);
    for (int ii = 0; ii < np; ++ii) {
        // P[ii][j] = feature_distance(db[j], db[pivot])  (lhs = gallery row j, rhs = pivot, ann.cpp:309)
        int st_ = launch_pair_distances(g->metric, nullptr, 1, g->dp, g->rows, g->dp, n, g->d, nullptr, (int)n, 1, row, s, cur);
        if (st_ != FIR_OK) return cleanup(st_);
        dem_step_kernel<<<nblk, RB, 0, s>>>(row, n, g->labels, cur, far_sum, ii < keep ? dm->P_raw + (size_t)ii * n : nullptr, blk_max, blk_arg, blk_min, 0, nullptr);
        dem_step_final_kernel<<<1, RB, 0, s>>>(blk_max, blk_arg, blk_min, nblk, ii, np, d_chain, d_min_other, cur, 0, nullptr);
    }
    if (cudaGetLastError() != cudaSuccess) return cleanup(fail(FIR_ERR_CUDA, "DEM build kernel launch failed"));
    dm->pivots.resize(np);
    dm->min_other.resize(np);
    cudaError_t e = cudaMemcpyAsync(dm->pivots.data(), d_chain, 4 * (size_t)np, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dm->min_other.data(), d_min_other, 4 * (size_t)np, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return cleanup(fail(FIR_ERR_CUDA, std::string("DEM build: ") + cudaGetErrorString(e)));
    for (int i = 0; i < np; ++i)
        if (dm->pivots[i] < 0) return cleanup(fail(FIR_ERR_UNSUPPORTED, "degenerate gallery: the farthest-point chain found no next pivot (the reference indexes dbImages[-1] here)"));
    // threshold: getThreshold(otherClassesDists, FAR) = element (int)(size*FAR) of the sorted list (ann.cpp:84-93), unless overridden (:277-279,340)
    if (params->threshold > 0) dm->threshold = params->threshold;
    else {
        std::vector<float> o(dm->min_other);
        int ind = (int)((float)o.size() * params->false_accept_rate);
        ind = std::max(0, std::min(ind, (int)o.size() - 1));
        std::nth_element(o.begin(), o.begin() + ind, o.end());
        dm->threshold = o[ind];
    }
    { int st2 = finalize_search_state(dm); if (st2 != FIR_OK) return cleanup(st2); }
    *out = dm;
    return cleanup(FIR_OK);
}

// Row-sharded build (collective: every rank of `c` calls it with its own shard handle).  The farthest-point chain is global:
// per step the current pivot's row + class are exchanged as raw bits (all-reduce(max) with zeros from the non-owners), every
// rank computes P[ii][j] for ITS rows j and its far-sum maximum / minimum other-class distance, the (maximum, row, minimum)
// triples are all-gathered and combined in the reference's scan order.  The next pivot never leaves the device: two small
// collectives per step, no host synchronisation inside the chain.  Pivots, threshold and the candidate order are those of
// ONE DirectedEnumeration over the whole gallery (ann.cpp:270-348).
int fir_shard_dem_build(fir_gallery* g, fir_comm* c, int64_t n_total, const fir_dem_params* params, fir_dem** out) {
    if (!out) return fail(FIR_ERR_BAD_ARG, "out is null");
    *out = nullptr;
    if (!g || !c || !params) return fail(FIR_ERR_BAD_ARG, "null argument");
    if (c->world > 1 && !nccl_api()) return fail(FIR_ERR_NCCL, "NCCL unavailable");
    FIR_CUDA_TRY(cudaSetDevice(g->device));
    const int64_t n = g->n, lo = g->index_offset, N = n_total;
    if (N < 2 || lo < 0 || lo + n > N) return fail(FIR_ERR_BAD_ARG, "shard rows [index_offset, index_offset + n) must lie inside n_total");
    int np = (int)((double)N * 0.015);                              // ann.cpp:373
    if (np < 5) np = 5;
    if (params->max_chain > 0 && params->max_chain < np) np = params->max_chain;
    if (np > N) np = (int)N;
    int keep = params->max_pivots > 0 ? std::min(params->max_pivots, 32) : 32;
    keep = std::min(keep, np);
    int pivot0 = params->pivot0;
    if (pivot0 < 0) {
        uint64_t z = (uint64_t)params->seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
        z ^= z >> 31; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 29;
        pivot0 = (int)(z % (uint64_t)N);
    }
    if (pivot0 >= N) return fail(FIR_ERR_BAD_ARG, "pivot0 out of range");
    fir_dem* dm = new fir_dem();
    dm->g = g; dm->chain_rows = np; dm->n_pivots = keep;
    dm->sharded = true; dm->comm = c; dm->row_lo = lo; dm->n_total = N;
    cudaStream_t s = g->stream;
    const int dp = g->dp, W = c->world;
    const int nblk = (int)std::min<int64_t>(ceil_div(n, RB), 1184);
    float* row = nullptr; double* far_sum = nullptr; double* blk_max = nullptr; int32_t* blk_arg = nullptr; float* blk_min = nullptr;
    int32_t* d_chain = nullptr; float* d_min_other = nullptr; int32_t* cur = nullptr; uint32_t* rec = nullptr; unsigned long long* trips = nullptr;
    auto cleanup = [&](int code) {
        cudaFree(row); cudaFree(far_sum); cudaFree(blk_max); cudaFree(blk_arg); cudaFree(blk_min); cudaFree(d_chain); cudaFree(d_min_other); cudaFree(cur);
        cudaFree(rec); cudaFree(trips);
        if (code != FIR_OK) fir_dem_destroy(dm);
        return code;
    };
    if (cudaMalloc(&row, 4 * (size_t)n) != cudaSuccess || cudaMalloc(&far_sum, 8 * (size_t)n) != cudaSuccess ||
        cudaMalloc(&blk_max, 8 * (size_t)nblk) != cudaSuccess || cudaMalloc(&blk_arg, 4 * (size_t)nblk) != cudaSuccess ||
        cudaMalloc(&blk_min, 4 * (size_t)nblk) != cudaSuccess || cudaMalloc(&d_chain, 4 * (size_t)np) != cudaSuccess ||
        cudaMalloc(&d_min_other, 4 * (size_t)np) != cudaSuccess || cudaMalloc(&cur, 4) != cudaSuccess ||
        cudaMalloc(&rec, 4 * (size_t)(dp + 1)) != cudaSuccess || cudaMalloc(&trips, 16 * (size_t)(W + 1)) != cudaSuccess ||
        cudaMalloc(&dm->P_raw, 4 * (size_t)keep * n) != cudaSuccess || cudaMalloc(&dm->P_search, 4 * (size_t)keep * n) != cudaSuccess ||
        cudaMalloc(&dm->in_tail, (size_t)n) != cudaSuccess || cudaMalloc(&dm->d_pivots, 4 * (size_t)keep) != cudaSuccess ||
        cudaMalloc(&dm->pivot_rows, 4 * (size_t)keep * dp) != cudaSuccess)
        return cleanup(fail(FIR_ERR_OOM, "sharded DEM build allocation failed"));
    cudaMemsetAsync(far_sum, 0, 8 * (size_t)n, s);
    cudaMemsetAsync(d_chain, 0xFF, 4 * (size_t)np, s);
    cudaMemcpyAsync(d_chain, &pivot0, 4, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(cur, &pivot0, 4, cudaMemcpyHostToDevice, s);
    NcclApi* NC = W > 1 ? nccl_api() : nullptr;
    unsigned long long* my_trip = trips + 2 * (size_t)W;            // this rank's triple; trips[0 .. 2W) receives everyone's
    {   // smallest shard of the communicator (decides, identically everywhere, whether the first round runs on tensor cores)
        unsigned long long v = (unsigned long long)n;
        cudaMemcpyAsync(my_trip, &v, 8, cudaMemcpyHostToDevice, s);
        if (NC) { ncclResult_t e = NC->AllReduce(my_trip, my_trip, 1, ncclUint64, ncclMin, c->comm, s);
                  if (e != ncclSuccess) return cleanup(fail(FIR_ERR_NCCL, std::string("ncclAllReduce (shard rows): ") + NC->GetErrorString(e))); }
        cudaMemcpyAsync(&v, my_trip, 8, cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) return cleanup(fail(FIR_ERR_CUDA, "sharded DEM build: stream error"));
        dm->min_shard_rows = (int64_t)v;
    }
    for (int ii = 0; ii < np; ++ii) {
        dem_pivot_record_kernel<<<1, 256, 0, s>>>(g->rows, dp, g->labels, lo, n, cur, rec);
        if (NC) { ncclResult_t e = NC->AllReduce(rec, rec, (size_t)dp + 1, ncclUint32, ncclMax, c->comm, s);
                  if (e != ncclSuccess) return cleanup(fail(FIR_ERR_NCCL, std::string("ncclAllReduce (pivot record): ") + NC->GetErrorString(e))); }
        if (ii < keep) cudaMemcpyAsync(dm->pivot_rows + (size_t)ii * dp, rec, 4 * (size_t)dp, cudaMemcpyDeviceToDevice, s);
        // P[ii][j] = feature_distance(db[j], pivot)  (lhs = gallery row j, rhs = the pivot, ann.cpp:309)
        int st_ = launch_pair_distances(g->metric, reinterpret_cast<const float*>(rec), 1, dp, g->rows, dp, n, g->d, nullptr, (int)n, 1, row, s);
        if (st_ != FIR_OK) return cleanup(st_);
        dem_step_kernel<<<nblk, RB, 0, s>>>(row, n, g->labels, cur, far_sum, ii < keep ? dm->P_raw + (size_t)ii * n : nullptr, blk_max, blk_arg, blk_min, lo,
                                            rec + dp);
        dem_step_final_kernel<<<1, RB, 0, s>>>(blk_max, blk_arg, blk_min, nblk, ii, np, d_chain, d_min_other, cur, lo, my_trip);
        if (NC) { ncclResult_t e = NC->AllGather(my_trip, trips, 2, ncclUint64, c->comm, s);
                  if (e != ncclSuccess) return cleanup(fail(FIR_ERR_NCCL, std::string("ncclAllGather (chain step): ") + NC->GetErrorString(e))); }
        else cudaMemcpyAsync(trips, my_trip, 16, cudaMemcpyDeviceToDevice, s);
        dem_pick_global_kernel<<<1, 32, 0, s>>>(trips, W, ii, np, d_chain, d_min_other, cur);
    }
    if (cudaGetLastError() != cudaSuccess) return cleanup(fail(FIR_ERR_CUDA, "sharded DEM build kernel launch failed"));
    dm->pivots.resize(np);
    dm->min_other.resize(np);
    cudaError_t e = cudaMemcpyAsync(dm->pivots.data(), d_chain, 4 * (size_t)np, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dm->min_other.data(), d_min_other, 4 * (size_t)np, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return cleanup(fail(FIR_ERR_CUDA, std::string("sharded DEM build: ") + cudaGetErrorString(e)));
    for (int i = 0; i < np; ++i)
        if (dm->pivots[i] < 0) return cleanup(fail(FIR_ERR_UNSUPPORTED, "degenerate gallery: the farthest-point chain found no next pivot"));
    if (params->threshold > 0) dm->threshold = params->threshold;
    else {
        std::vector<float> o(dm->min_other);
        int ind = (int)((float)o.size() * params->false_accept_rate);
        ind = std::max(0, std::min(ind, (int)o.size() - 1));
        std::nth_element(o.begin(), o.begin() + ind, o.end());
        dm->threshold = o[ind];
    }
    { int st2 = finalize_search_state(dm); if (st2 != FIR_OK) return cleanup(st2); }
    *out = dm;
    return cleanup(FIR_OK);
}

int fir_dem_from_state(fir_gallery* g, const int32_t* pivots, int32_t n_pivots, const float* P, float threshold, fir_dem** out) {
    if (!out) return fail(FIR_ERR_BAD_ARG, "out is null");
    *out = nullptr;
    if (!g || !pivots || !P || n_pivots < 1 || n_pivots > 32 || n_pivots > g->n) return fail(FIR_ERR_BAD_ARG, "bad DEM state");
    for (int i = 0; i < n_pivots; ++i)
        if (pivots[i] < 0 || pivots[i] >= g->n) return fail(FIR_ERR_BAD_ARG, "pivot out of range");
    FIR_CUDA_TRY(cudaSetDevice(g->device));
    fir_dem* dm = new fir_dem();
    dm->g = g; dm->n_pivots = n_pivots; dm->chain_rows = n_pivots; dm->threshold = threshold;
    dm->pivots.assign(pivots, pivots + n_pivots);
    dm->min_other.assign((size_t)n_pivots, 0.f);
    const int64_t n = g->n;
    if (cudaMalloc(&dm->P_raw, 4 * (size_t)n_pivots * n) != cudaSuccess || cudaMalloc(&dm->P_search, 4 * (size_t)n_pivots * n) != cudaSuccess ||
        cudaMalloc(&dm->in_tail, (size_t)n) != cudaSuccess || cudaMalloc(&dm->d_pivots, 4 * (size_t)n_pivots) != cudaSuccess) {
        fir_dem_destroy(dm);
        return fail(FIR_ERR_OOM, "DEM state allocation failed");
    }
    cudaError_t e = cudaMemcpy(dm->P_raw, P, 4 * (size_t)n_pivots * n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { fir_dem_destroy(dm); return fail(FIR_ERR_CUDA, cudaGetErrorString(e)); }
    int st = finalize_search_state(dm);
    if (st != FIR_OK) { fir_dem_destroy(dm); return st; }
    *out = dm;
    return FIR_OK;
}

int fir_dem_destroy(fir_dem* dm) {
    if (!dm) return FIR_OK;
    cudaFree(dm->P_raw); cudaFree(dm->P_search); cudaFree(dm->in_tail); cudaFree(dm->d_pivots);
    if (dm->pt) fir_gallery_destroy(dm->pt);
    cudaFree(dm->center); cudaFree(dm->exclude); cudaFree(dm->sp_rows); cudaFree(dm->sp_mask); cudaFree(dm->sp_tail);
    cudaFree(dm->pivot_rows);
    delete dm;
    return FIR_OK;
}

int fir_dem_info(const fir_dem* dm, int32_t* n_pivots, int32_t* chain_rows, float* threshold) {
    if (!dm) return fail(FIR_ERR_BAD_ARG, "dem is null");
    if (n_pivots) *n_pivots = dm->n_pivots;
    if (chain_rows) *chain_rows = dm->chain_rows;
    if (threshold) *threshold = dm->threshold;
    return FIR_OK;
}
int fir_dem_search_stats(fir_dem* dm, int32_t* gpu_launches, int32_t* tensor_round, double* candidates_kernel_ms, int32_t* candidates_kernel_launches) {
    if (!dm) return fail(FIR_ERR_BAD_ARG, "dem is null");
    if (gpu_launches) *gpu_launches = dm->last_launches;
    if (tensor_round) *tensor_round = dm->pt ? 1 : 0;
    if (candidates_kernel_ms) *candidates_kernel_ms = 0;
    if (candidates_kernel_launches) *candidates_kernel_launches = 0;
    if (dm->pt && dm->pt->profiling) return fir_profile_read(dm->pt, FIR_KERNEL_L2_CANDIDATES, candidates_kernel_ms, candidates_kernel_launches);
    return FIR_OK;
}
int fir_dem_get_pivots(const fir_dem* dm, int32_t* out_pivots) {
    if (!dm || !out_pivots) return fail(FIR_ERR_BAD_ARG, "null argument");
    std::memcpy(out_pivots, dm->pivots.data(), 4 * (size_t)dm->n_pivots);
    return FIR_OK;
}
int fir_dem_get_pivot_matrix(const fir_dem* dm, float* out_P) {
    if (!dm || !out_P) return fail(FIR_ERR_BAD_ARG, "null argument");
    FIR_CUDA_TRY(cudaMemcpy(out_P, dm->P_raw, 4 * (size_t)dm->n_pivots * dm->g->n, cudaMemcpyDeviceToHost));
    return FIR_OK;
}
int fir_dem_get_min_other(const fir_dem* dm, float* out) {
    if (!dm || !out) return fail(FIR_ERR_BAD_ARG, "null argument");
    std::memcpy(out, dm->min_other.data(), 4 * (size_t)dm->chain_rows);
    return FIR_OK;
}

int fir_dem_search(fir_dem* dm, const float* queries, int64_t nq, int32_t count_to_check, int32_t memspace, int32_t* out_idx,
                   float* out_dist, uint8_t* out_below, int32_t* out_evals) {
    if (!dm) return fail(FIR_ERR_BAD_ARG, "dem is null");
    if (nq < 0 || (nq > 0 && (!queries || !out_idx))) return fail(FIR_ERR_BAD_ARG, "bad arguments");
    if (nq == 0) return FIR_OK;
    fir_gallery* g = dm->g;
    FIR_CUDA_TRY(cudaSetDevice(g->device));
    cudaStream_t s = g->stream;
    const int64_t n = g->n;
    const int S = dm->n_pivots;
    const bool sharded = dm->sharded && dm->comm && dm->comm->world > 1;
    const int64_t n_all = dm->sharded ? dm->n_total : n;                                     // rows of the WHOLE gallery
    const int32_t ioff = dm->sharded ? (int32_t)dm->row_lo : 0;                               // row shards: QState rows are global
    const int M = (count_to_check > 0 && count_to_check < n_all) ? count_to_check : (int)n_all;   // ann.h:20-22
    // (the reference has undefined behaviour for M < S — partial_sort with middle < first, ann.cpp:469; here the
    //  candidate phase simply does not run, which is also what its while-loop at :472 does.)
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    // (row shards: chunking must be identical on every rank — the rounds are collective — so it is sized from the largest shard)
    const int64_t n_sz = sharded ? ceil_div(n_all, dm->comm->world) : n;
    const int QC = (int)std::max<int64_t>(8, std::min<int64_t>(std::min<int64_t>(nq, 2048), ((int64_t)4 << 30) / (4 * n_sz)));   // lik chunk ≤ 4 GiB
    int cap_max = (int)std::min<int64_t>(n_sz, std::max<int64_t>(256, ((int64_t)256 << 20) / ((int64_t)QC * 12)));
    if (sharded) cap_max = std::min(cap_max, 1024);          // the exchange sorts world x cap records per query in shared memory
    size_t need = al(4 * (size_t)nq * g->dp) + 2 * al(4 * (size_t)nq * S) + al(sizeof(QState) * (size_t)nq) + al(4 * (size_t)nq) +
                  al(4 * (size_t)QC * n) + al(4 * (size_t)QC * 2048) + 3 * al(4 * (size_t)QC * cap_max) + 14 * al(4 * (size_t)QC) + al(4 * (size_t)QC * TIE_CAP) +
                  al((size_t)nq * 9) + al(4 * (size_t)nq) * 2 + 65536;
    static const int fast_on = [] { const char* e = getenv("FIR_DEM_FAST"); return e ? atoi(e) : 1; }();
    const int64_t ng = ceil_div(n, 32);
    const bool fast = fast_on != 0 && ng >= 1024 && !sharded;   // first-round select through per-warp minima (dem_fast_*)
    if (fast) need += al(4 * (size_t)QC * ng) + 2 * al(4 * (size_t)QC * FAST_CAP) + 4 * al(4 * (size_t)QC);
    const bool tensor0 = dm->pt != nullptr && M > S;
    const int k0 = tensor0 ? std::min(kRound0K, M - S) : 0;
    if (tensor0) need += 5 * al(4 * (size_t)nq * k0) + 3 * al(4 * (size_t)nq);
    FIR_TRY(g->ws.reserve(need));
    int launches = 0;
    // queries → zero-padded device rows
    const float* dq = nullptr;
    if (memspace == FIR_DEVICE && g->d == g->dp) dq = queries;
    else {
        float* buf = (float*)g->ws.take(4 * (size_t)nq * g->dp);
        if (!buf) return fail(FIR_ERR_INTERNAL, "workspace underestimated (dem queries)");
        if (memspace == FIR_HOST) {
            if (g->d != g->dp) FIR_CUDA_TRY(cudaMemsetAsync(buf, 0, 4 * (size_t)nq * g->dp, s));
            FIR_CUDA_TRY(cudaMemcpy2DAsync(buf, 4 * (size_t)g->dp, queries, 4 * (size_t)g->d, 4 * (size_t)g->d, (size_t)nq, cudaMemcpyHostToDevice, s));
        } else FIR_TRY(launch_pad_rows(queries, nq, g->d, buf, g->dp, s));
        dq = buf;
    }
    int32_t* pcand = (int32_t*)g->ws.take(4 * (size_t)nq * S);
    float* pd = (float*)g->ws.take(4 * (size_t)nq * S);
    QState* st = (QState*)g->ws.take(sizeof(QState) * (size_t)nq);
    int32_t* active = (int32_t*)g->ws.take(4 * (size_t)nq);
    float* lik = (float*)g->ws.take(4 * (size_t)QC * n);
    uint32_t* hist = (uint32_t*)g->ws.take(4 * (size_t)QC * 2048);
    int32_t* cand = (int32_t*)g->ws.take(4 * (size_t)QC * cap_max);
    uint32_t* cand_key = (uint32_t*)g->ws.take(4 * (size_t)QC * cap_max);
    float* cdist = (float*)g->ws.take(4 * (size_t)QC * cap_max);
    int32_t* want = (int32_t*)g->ws.take(4 * (size_t)QC);
    uint32_t* prefix = (uint32_t*)g->ws.take(4 * (size_t)QC);
    int32_t* want_rem = (int32_t*)g->ws.take(4 * (size_t)QC);
    int32_t* take_all = (int32_t*)g->ws.take(4 * (size_t)QC);
    int32_t* cnt = (int32_t*)g->ws.take(4 * (size_t)QC);
    int32_t* tie_cnt = (int32_t*)g->ws.take(4 * (size_t)QC);
    int32_t* overflow = (int32_t*)g->ws.take(4 * (size_t)QC);
    int32_t* tie_buf = (int32_t*)g->ws.take(4 * (size_t)QC * TIE_CAP);
    uint32_t* key_sel = (uint32_t*)g->ws.take(4 * (size_t)QC);
    int32_t* tie_take = (int32_t*)g->ws.take(4 * (size_t)QC);
    int32_t* round_n = (int32_t*)g->ws.take(4 * (size_t)QC);
    int32_t* last_tie = (int32_t*)g->ws.take(4 * (size_t)QC);
    int32_t* counters = (int32_t*)g->ws.take(64);
    float* gmin = nullptr; uint32_t* fkey = nullptr; int32_t* frow = nullptr; int32_t* fcnt = nullptr; int32_t* fast_ok = nullptr;
    uint32_t* t_key = nullptr; int32_t* t_all = nullptr;
    if (fast) {
        gmin = (float*)g->ws.take(4 * (size_t)QC * ng);
        fkey = (uint32_t*)g->ws.take(4 * (size_t)QC * FAST_CAP);
        frow = (int32_t*)g->ws.take(4 * (size_t)QC * FAST_CAP);
        fcnt = (int32_t*)g->ws.take(4 * (size_t)QC);
        fast_ok = (int32_t*)g->ws.take(4 * (size_t)QC);
        t_key = (uint32_t*)g->ws.take(4 * (size_t)QC);
        t_all = (int32_t*)g->ws.take(4 * (size_t)QC);
        if (!gmin || !fkey || !frow || !fcnt || !fast_ok || !t_key || !t_all) return fail(FIR_ERR_INTERNAL, "workspace underestimated (dem fast round)");
        FIR_CUDA_TRY(cudaFuncSetAttribute(dem_fast_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    }
    int32_t* o_idx = out_idx; float* o_dist = out_dist; uint8_t* o_below = out_below; int32_t* o_evals = out_evals;
    if (memspace == FIR_HOST) {
        o_idx = (int32_t*)g->ws.take(4 * (size_t)nq);
        o_dist = (float*)g->ws.take(4 * (size_t)nq);
        o_evals = (int32_t*)g->ws.take(4 * (size_t)nq);
        o_below = (uint8_t*)g->ws.take((size_t)nq);
    }
    if (!pcand || !pd || !st || !active || !lik || !hist || !cand || !cand_key || !cdist || !want || !prefix || !want_rem || !take_all || !cnt || !tie_cnt || !overflow || !tie_buf || !key_sel ||
        !tie_take || !round_n || !last_tie || !counters || !o_idx || (memspace == FIR_HOST && (!o_dist || !o_evals || !o_below)))
        return fail(FIR_ERR_INTERNAL, "workspace underestimated (dem search)");

    // 1. pivot distances + pivot phase
    if (dm->sharded) {                                      // the pivots' rows are replicated on every rank: an S-row gallery, identity candidates
        FIR_TRY(launch_pair_distances(g->metric, dq, nq, g->dp, dm->pivot_rows, g->dp, S, g->d, nullptr, S, 0, pd, s));
    } else {
        dem_fill_pivot_cand_kernel<<<(unsigned)ceil_div(nq * S, 256), 256, 0, s>>>(dm->d_pivots, S, nq, pcand);
        FIR_TRY(launch_pair_distances(g->metric, dq, nq, g->dp, g->rows, g->dp, n, g->d, pcand, S, 0, pd, s));
    }
    dem_pivot_phase_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, s>>>(pd, nq, S, dm->d_pivots, dm->threshold, M, st);
    launches += 3;
    FIR_CUDA_TRY(cudaMemsetAsync(counters, 0, 64, s));
    // row shards: pack this rank's round → all-gather → the same in-order fold on every rank (dem_global_reduce_kernel)
    auto exchange_reduce = [&](const int32_t* c_, const uint32_t* ck_, const float* cd_, const int32_t* rn_, const int32_t* ql_, int nqc_, int cap_,
                               int round_size_, int32_t* n_active_) -> int {
        fir_comm* c = dm->comm;
        const size_t cells = (size_t)nqc_ * cap_;
        uint32_t* rec = nullptr; uint32_t* rec_all = nullptr;
        FIR_TRY(comm_take(c, 10, 12 * cells, s, (void**)&rec));
        FIR_TRY(comm_take(c, 11, 12 * cells * c->world, s, (void**)&rec_all));
        dem_pack_records_kernel<<<(unsigned)ceil_div((int64_t)cells, 256), 256, 0, s>>>(c_, ck_, cd_, rn_, nqc_, cap_, ioff, rec);
        FIR_NCCL_TRY(nccl_api()->AllGather(rec, rec_all, 3 * cells, ncclUint32, c->comm, s));
        int P2 = 1;
        while (P2 < c->world * cap_) P2 <<= 1;
        const size_t smem = (size_t)P2 * 12;
        FIR_CUDA_TRY(cudaFuncSetAttribute(dem_global_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 49152)));
        dem_global_reduce_kernel<<<(unsigned)nqc_, 256, smem, s>>>(rec_all, c->world, nqc_, cap_, ql_, round_size_, dm->threshold, M, dm->tail_size, S, st, n_active_);
        FIR_CUDA_TRY(cudaGetLastError());
        launches += 2;
        return FIR_OK;
    };
    // 1b. round 0: the first k0 candidates of every query through the tensor-core brute force in pivot space
    if (tensor0) {
        int32_t* t_idx = (int32_t*)g->ws.take(4 * (size_t)nq * k0);
        float* t_dist = (float*)g->ws.take(4 * (size_t)nq * k0);
        int32_t* c0 = (int32_t*)g->ws.take(4 * (size_t)nq * k0);
        uint32_t* ck0 = (uint32_t*)g->ws.take(4 * (size_t)nq * k0);
        float* cd0 = (float*)g->ws.take(4 * (size_t)nq * k0);
        int32_t* rn0 = (int32_t*)g->ws.take(4 * (size_t)nq);
        uint32_t* ks0 = (uint32_t*)g->ws.take(4 * (size_t)nq);
        int32_t* lt0 = (int32_t*)g->ws.take(4 * (size_t)nq);
        if (!t_idx || !t_dist || !c0 || !ck0 || !cd0 || !rn0 || !ks0 || !lt0) return fail(FIR_ERR_INTERNAL, "workspace underestimated (dem round 0)");
        dm->pt->stream = s; dm->pt->ws.stream = s;
        if (dm->pt->profiling != g->profiling) { dm->pt->profiling = g->profiling; dm->pt->ev_used = 0; }   // follows fir_profile_enable of the gallery
        dm->pt->stats = fir_search_stats{};
        FIR_TRY(tensor_search_topk(dm->pt, pd, nq, k0, FIR_DEVICE, t_idx, t_dist));
        launches += dm->pt->stats.gpu_launches;
        dem_round0_merge_kernel<<<(unsigned)ceil_div(nq, 4), 128, 0, s>>>(t_idx, t_dist, k0, nq, pd, S, dm->P_raw, n, dm->sp_rows, dm->sp_mask, dm->sp_tail,
                                                                         dm->n_special, M, st, c0, ck0, rn0, ks0, lt0);
        FIR_TRY(launch_pair_distances(g->metric, dq, nq, g->dp, g->rows, g->dp, n, g->d, c0, k0, 0, cd0, s));
        if (sharded) FIR_TRY(exchange_reduce(c0, ck0, cd0, rn0, nullptr, (int)nq, k0, k0, counters + 2));
        else dem_reduce_kernel<<<(unsigned)ceil_div(nq, 4), 128, 0, s>>>(cd0, c0, ck0, nullptr, (int)nq, k0, rn0, ks0, lt0, dm->threshold, M, dm->tail_size, S, st, counters + 2);
        FIR_CUDA_TRY(cudaGetLastError());
        launches += 3;
    }
    // 2. queries that go on to the candidate walk
    dem_active_list_kernel<<<1, 1024, 0, s>>>(st, nq, active, counters);
    int32_t n_act = 0;
    FIR_CUDA_TRY(cudaMemcpyAsync(&n_act, counters, 4, cudaMemcpyDeviceToHost, s));
    FIR_CUDA_TRY(cudaStreamSynchronize(s));
    for (int lo = 0; lo < n_act; lo += QC) {
        const int nqc = std::min(QC, n_act - lo);
        const int32_t* ql = active + lo;
        { auto* ev = g->prof_begin(FIR_KERNEL_DEM_LIKELIHOOD);
          if (nqc >= 24) dem_likelihood_kernel<32><<<dim3((unsigned)ceil_div(n, 256), (unsigned)ceil_div(nqc, 32)), 256, 0, s>>>(pd, ql, nqc, S, dm->P_search, n, dm->in_tail, lik, gmin, ng);
          else dem_likelihood_kernel<8><<<dim3((unsigned)ceil_div(n, 256), (unsigned)ceil_div(nqc, 8)), 256, 0, s>>>(pd, ql, nqc, S, dm->P_search, n, dm->in_tail, lik, gmin, ng);
          g->prof_end(ev); }
        int round_size = 256;
        int remaining = nqc;
        bool first_round = !tensor0;                         // after round 0 the walk resumes behind its closing key: the general select
        while (remaining > 0) {
            const int cap = (int)std::min<int64_t>(round_size, cap_max);
            dem_want_kernel<<<(unsigned)ceil_div(nqc, 128), 128, 0, s>>>(st, ql, nqc, cap, M, want);
            const unsigned hb = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256 * 16), 64));
            const int32_t* skip = nullptr;
            if (fast && first_round) {
                // bound from the per-warp minima (the same radix select over an array 32x smaller), one collecting pass, sort
                const unsigned hg = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(ng, 256 * 16), 64));
                for (int pass = 0; pass < 3; ++pass) {
                    FIR_CUDA_TRY(cudaMemsetAsync(hist, 0, 4 * (size_t)nqc * 2048, s));
                    dem_hist_kernel<<<dim3(hg, (unsigned)nqc), 256, 0, s>>>(gmin, ql, nqc, ng, st, pass, prefix, take_all, hist, nullptr);
                    dem_pick_kernel<<<(unsigned)nqc, 32, 0, s>>>(hist, ql, nqc, st, pass, want, prefix, want_rem, take_all, key_sel, tie_take, round_n, nullptr);
                }
                FIR_CUDA_TRY(cudaMemcpyAsync(t_key, key_sel, 4 * (size_t)nqc, cudaMemcpyDeviceToDevice, s));
                FIR_CUDA_TRY(cudaMemcpyAsync(t_all, take_all, 4 * (size_t)nqc, cudaMemcpyDeviceToDevice, s));
                FIR_CUDA_TRY(cudaMemsetAsync(fcnt, 0, 4 * (size_t)nqc, s));
                dem_fast_collect_kernel<<<dim3(hb, (unsigned)nqc), 256, 0, s>>>(lik, ql, nqc, n, st, t_all, t_key, fkey, frow, fcnt);
                dem_fast_finalize_kernel<<<(unsigned)nqc, 256, 16384 * 8, s>>>(ql, nqc, st, t_all, want, fkey, frow, fcnt, cap, cand, cand_key, round_n, key_sel, last_tie,
                                                                              take_all, fast_ok);
                skip = fast_ok;
            }
            first_round = false;
            for (int pass = 0; pass < 3; ++pass) {
                FIR_CUDA_TRY(cudaMemsetAsync(hist, 0, 4 * (size_t)nqc * 2048, s));
                dem_hist_kernel<<<dim3(hb, (unsigned)nqc), 256, 0, s>>>(lik, ql, nqc, n, st, pass, prefix, take_all, hist, skip, ioff);
                dem_pick_kernel<<<(unsigned)nqc, 32, 0, s>>>(hist, ql, nqc, st, pass, want, prefix, want_rem, take_all, key_sel, tie_take, round_n, skip);
            }
            FIR_CUDA_TRY(cudaMemsetAsync(cnt, 0, 4 * (size_t)nqc, s));
            FIR_CUDA_TRY(cudaMemsetAsync(tie_cnt, 0, 4 * (size_t)nqc, s));
            dem_collect_fast_kernel<<<dim3(hb, (unsigned)nqc), 256, 0, s>>>(lik, ql, nqc, n, st, take_all, key_sel, cap, cand, cand_key, cnt, tie_buf, tie_cnt, skip, ioff);
            dem_collect_fixup_kernel<<<(unsigned)ceil_div(nqc, 128), 128, 0, s>>>(ql, nqc, st, take_all, key_sel, tie_take, cap, cand, cand_key, cnt, tie_buf, tie_cnt,
                                                                               last_tie, overflow, skip);
            dem_collect_kernel<<<(unsigned)ceil_div(nqc, 4), 128, 0, s>>>(lik, ql, nqc, n, st, overflow, key_sel, tie_take, cap, cand, cand_key, last_tie, ioff);
            // exact distances of the round's candidates: queries are addressed through the active list
            FIR_TRY(launch_pair_distances(g->metric, dq, nqc, g->dp, g->rows, g->dp, n, g->d, cand, cap, 0, cdist, s, nullptr, ql));
            FIR_CUDA_TRY(cudaMemsetAsync(counters + 1, 0, 4, s));
            if (sharded) FIR_TRY(exchange_reduce(cand, cand_key, cdist, round_n, ql, nqc, cap, cap, counters + 1));
            else dem_reduce_kernel<<<(unsigned)ceil_div(nqc, 4), 128, 0, s>>>(cdist, cand, cand_key, ql, nqc, cap, round_n, key_sel, last_tie, dm->threshold, M,
                                                                       dm->tail_size, S, st, counters + 1);
            FIR_CUDA_TRY(cudaMemcpyAsync(&remaining, counters + 1, 4, cudaMemcpyDeviceToHost, s));
            FIR_CUDA_TRY(cudaStreamSynchronize(s));
            if (round_size < cap_max) round_size = (int)std::min<int64_t>((int64_t)round_size * 8, cap_max);
        }
    }
    dem_output_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, s>>>(st, nq, dm->sharded ? 0 : g->index_offset, o_idx, o_dist, o_below, o_evals);
    FIR_CUDA_TRY(cudaGetLastError());
    dm->last_launches = launches + 2;
    if (memspace == FIR_HOST) {
        FIR_CUDA_TRY(cudaMemcpyAsync(out_idx, o_idx, 4 * (size_t)nq, cudaMemcpyDeviceToHost, s));
        if (out_dist) FIR_CUDA_TRY(cudaMemcpyAsync(out_dist, o_dist, 4 * (size_t)nq, cudaMemcpyDeviceToHost, s));
        if (out_below) FIR_CUDA_TRY(cudaMemcpyAsync(out_below, o_below, (size_t)nq, cudaMemcpyDeviceToHost, s));
        if (out_evals) FIR_CUDA_TRY(cudaMemcpyAsync(out_evals, o_evals, 4 * (size_t)nq, cudaMemcpyDeviceToHost, s));
        FIR_CUDA_TRY(cudaStreamSynchronize(s));
    }
    return FIR_OK;
}

}  // extern "C"
