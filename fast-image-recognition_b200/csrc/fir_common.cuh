// fir_common.cuh — shared declarations for libfir_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/fir_b200.h"

namespace fir {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define FIR_CUDA_TRY(expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            return ::fir::fail(_e == cudaErrorMemoryAllocation ? FIR_ERR_OOM : FIR_ERR_CUDA,       \
                               std::string(#expr) + ": " + cudaGetErrorString(_e));                \
        }                                                                                           \
    } while (0)

#define FIR_TRY(expr)                 \
    do {                              \
        int _s = (expr);              \
        if (_s != FIR_OK) return _s;  \
    } while (0)

// Growable device scratch, bump-allocated per call (256-byte aligned).
struct Workspace {
    char* base = nullptr;
    size_t cap = 0, off = 0;
    uint64_t generation = 0;     // bumped by every reserve(): pointers handed out before it are dead
    cudaStream_t stream = 0;
    int reserve(size_t bytes);   // make sure cap >= bytes (may sync + realloc); resets off
    void* take(size_t bytes);    // nullptr if it does not fit
    void release();
};

constexpr int kExactTile = 64;   // queries x gallery rows per block tile of the exact kernel
constexpr int kExactChunk = 32;  // dims staged per cp.async stage
constexpr int kRowPad = 32;      // fp32 rows are stored zero-padded to a multiple of this

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// exact tile kernel modes
enum { MODE_TOPK = 0, MODE_CLASSMIN = 1, MODE_PNN = 2 };

struct ExactParams {
    const float* q;   int64_t nq;  int ldq;
    const float* x;   int64_t n;   int ldx;
    const int32_t* labels;
    int d_end;                       // dims used (max_features or D); the mean divides by this
    int k;                           // MODE_TOPK list length
    int mode;
    int nsplit; int64_t tiles_per_split;
    float* part_dist; int32_t* part_idx;       // [nq][nsplit][k], local row indices
    unsigned long long* cls_key;               // [nq][C] packed (ordered dist bits, local idx)
    double* cls_score;                         // [nq][C]
    int n_classes; double two_var;
    const int32_t* qmap; const int32_t* n_active;   // optional query indirection (tensor-path fallback)
    int64_t active_offset; int64_t active_cap;      // this launch serves active positions [offset, offset+cap) (cap 0 = all)
};

int launch_exact_tiles(int metric, const ExactParams& p, cudaStream_t s);
int launch_merge_parts(const float* pd, const int32_t* pi, int n_parts, int64_t part_stride, int64_t q_stride,
                       int64_t nq, int k, int64_t index_offset, const int32_t* qmap, const int32_t* n_active,
                       float* od, int32_t* oi, cudaStream_t s, int64_t active_offset = 0, int64_t active_cap = 0);
int launch_fill_u64(unsigned long long* p, int64_t n, unsigned long long v, cudaStream_t s);
int launch_classmin_finalize(const unsigned long long* keys, int64_t n, int64_t index_offset, float* omin, int32_t* oarg, cudaStream_t s);
int launch_pnn_finalize(double* scores, int64_t nq, int n_classes, double n_total, int32_t* olabel, cudaStream_t s);
int launch_pair_distances(int metric, const float* q, int64_t nq, int ldq, const float* x, int ldx, int64_t n, int d_end,
                          const int32_t* cand, int r, int gallery_is_lhs, float* out, cudaStream_t s, const int32_t* qsel = nullptr,
                          const int32_t* qmap = nullptr);
int launch_normalize_rows(float* rows, int64_t n, int d, int ld, int metric, cudaStream_t s);
int launch_pad_rows(const float* src, int64_t n, int d, float* dst, int ld, cudaStream_t s);

// ---- small-batch streaming path: stream_kernels.cu --------------------------------------------------
int launch_stream_distances(int metric, int nq, const float* q, int ldq, const float* x, int ldx, int64_t n, int d_end, float* out,
                            int64_t out_stride, int n_sm, cudaStream_t s);
int launch_stream_topk(const float* dist, int64_t stride, int64_t n, int nq, int k, int nseg, float* part_d, int32_t* part_i, cudaStream_t s);
int launch_stream_class(const float* dist, int64_t stride, int64_t n, int nq, const int32_t* labels, int n_classes, int mode, double two_var,
                        unsigned long long* cls_key, double* cls_score, cudaStream_t s);
constexpr int kStreamMaxQueries = 8;
constexpr int kStreamMaxK = 16;

// ---- tensor-core (tcgen05) L2 candidate path: l2_tensor.cu -------------------------------------
struct TensorSide {              // fp16 shadow of a set of fp32 rows
    __half* h = nullptr;         // [rows_padded][dph], scaled by `scale`
    float* norm2 = nullptr;      // ||x||^2 (fp32 of the fp64 sum), +inf for padding rows
    float* resid = nullptr;      // ||x - h/scale|| rounded up
    float* row_scale = nullptr;  // per-row power-of-two scale (query side); the gallery side uses one global scale in meta[1]
    unsigned int* meta = nullptr; // [0] = bits of max|x|, [1] = bits of the power-of-two scale (device)
    int64_t rows = 0, rows_padded = 0;
    int64_t perm_a = 1, perm_b = 0;   // shadow row p = original row (p*perm_a + perm_b) mod rows
    int dph = 0;
    bool folded = false;         // gallery side: columns d, d+1 hold ‖x‖²·scale·2^a as an fp16 hi/lo pair, meta[2] = a (see fold_norm_kernel)
};

size_t tensor_side_bytes(int64_t rows, int d, int row_tile);
int tensor_fold_norms(TensorSide* side, int d, const float* d_stats, cudaStream_t s);     // gallery side, after tensor_pack_side
int tensor_pack_side(const float* rows, int64_t n, int ld, int d, int row_tile, void* buf, TensorSide* out,
                     float* d_stats /*[2]: max ||x||, max resid (device)*/, bool permute, bool per_row_scale, cudaStream_t s,
                     const float* center = nullptr /*[d] device: shadow of x - center*/, const unsigned char* exclude = nullptr /*[n] device*/);
struct TensorSearchArgs {
    const TensorSide* gal; const TensorSide* qry;
    const CUtensorMap* tmap_a; const CUtensorMap* tmap_b;
    int d; int R;                 // candidates kept per (query, slot)
    int n_slots;
    float* cand_val; int32_t* cand_idx;   // [nq][n_slots][R]
    float* slot_bound;                    // [nq][n_slots]
    const float* seed_thr;                // optional [nq] seed thresholds (approximate squared distances)
    const int32_t* skip_if_zero;          // optional device counter: the launch does nothing when it reads 0
    int mins_only;                        // seed pass: per-slot column-group minima instead of top-R lists (R must be 4)
    int grid; int n_sm;
    int ctas;                     // 1: cta_group::1 kernel, 2: CTA-pair kernel (grid counts pairs)
    int nf;                       // the gallery shadow carries its row norms as two extra columns and the queries are packed to match (see fold_norm_kernel)
    int nkb;                      // k-blocks to contract (0 = all of the shadow's; fewer = a prefix of the dimensions)
    int64_t row_bytes;            // bytes of one shadow row for the partition's L2 test (0 = gal->dph * 2)
    unsigned int* sync_ctr;       // optional zeroed device word: lets the pairs of a full round re-align (galleries larger than L2)
};
int tensor_plan(int64_t nq, int64_t n, int n_sm, int ctas, int* grid, int* n_slots, int* min_slots = nullptr, int64_t row_bytes = 0);
int tensor_cta_mode();
int tensor_encode_map(CUtensorMap* map, const __half* base, int64_t rows_padded, int dph, int box_rows);
int launch_tensor_candidates(const TensorSearchArgs& a, cudaStream_t s);

bool tensor_path_supported(int d);

// How far an approximate candidate value can be from the reference's own result, per query:
//   |approx − dist_scale·reference| ≤ E + rel·(dist_scale·reference)
struct ErrModel {
    int kind;                 // 0: fp16 tensor-core L2 (Cauchy–Schwarz on measured residuals); 1: relative only; 2: E from L1 norms;
                              // 3: KL entropy form: E = (‖q‖₁ + max‖x‖₁)·(abs_coef + lam_coef·Λ), Λ from the smallest positive elements
    int d, nkb;
    const float* q_norm2; const float* q_resid; const float* gal_stats;     // kind 0
    double rel;               // relative part (kind 0: the reference's sequential-sum error (D+4)·2⁻²⁴)
    double abs_coef;          // kind 2: E = abs_coef · (‖q‖₁ + max‖x‖₁)
    const float* q_l1; const float* x_l1_max;                               // kind 2, 3 (device)
    double lam_coef; const float* q_minpos; const float* x_minpos;          // kind 3 (device)
    double extra_nx2;         // kind 0: added to E as extra_nx2 · (max ‖x‖)² (the folded-norm representation error)
    double dist_scale;        // approx units per reference-distance unit: D on the tensor path (squared distance), 1 otherwise
};
int launch_prune(float* cand_val, int32_t* cand_idx, int64_t nq, int rt, int k, const ErrModel& em, cudaStream_t s, const int32_t* n_active = nullptr,
                 uint32_t* pair_cells = nullptr, int32_t* pair_count = nullptr);
// exact distances of the listed (query, candidate) cells: cell = q * rt + position in the query's candidate list
int launch_pair_list(int metric, const float* q, int ldq, const float* x, int ldx, int d_end, const uint32_t* cells, const int32_t* count,
                     int64_t max_cells, int rt, const int32_t* cand_idx, float* out, cudaStream_t s);
int launch_select(const float* cand_exact, const int32_t* cand_idx, const float* slot_bound, int64_t nq, int n_slots, int R, int k,
                  const ErrModel& em, int64_t index_offset, float* out_dist, int32_t* out_idx, int32_t* flagged, int32_t* n_flagged,
                  unsigned char* fail_flags, float* max_bound, cudaStream_t s, const int32_t* n_active = nullptr, float* seed_out = nullptr);

}  // namespace fir

// device-side helpers ------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace fir {

__device__ __forceinline__ uint32_t ordered_bits(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(b);
}

// glibc 2.39 logf, __logf_fma build (see oracle/fir_oracle.c for provenance of the constants).
__device__ const double kLogfTab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2},
};

__device__ __forceinline__ float glibc_logf(float x) {
    uint32_t ix = __float_as_uint(x);
    // (glibc's `x == 1 → +0` shortcut is redundant in round-to-nearest: table entry 9 is {1, 0}, so the main path gives +0)
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return -__int_as_float(0x7f800000);
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return __int_as_float(0x7fc00000);
        ix = __float_as_uint(__fmul_rn(x, 8388608.0f));
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> 19) & 15u);
    int k = (int)tmp >> 23;
    uint32_t iz = ix - (tmp & 0xff800000u);
    double invc = kLogfTab[i][0], logc = kLogfTab[i][1];
    double z = (double)__uint_as_float(iz);
    double r = __fma_rn(z, invc, -1.0);
    double y0 = __fma_rn((double)k, 0x1.62e42fefa39efp-1, logc);
    double r2 = __dmul_rn(r, r);
    double y = __fma_rn(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
    y = __fma_rn(-0x1.00ea348b88334p-2, r2, y);
    y = __fma_rn(y, r2, __dadd_rn(y0, r));
    return __double2float_rn(y);
}

// One step of feature_distance (qt_cpp/db_features.cpp:26,29-36): fp32, no FMA contraction,
// IEEE round-to-nearest for every operation, in the reference's operation order.
// The chi-square and KL steps are written branch-free: the reference's `if (l + r > 0)` / `if (l > 0)` guards become
// selects around the same arithmetic on a safe operand (about half of ReLU-style features are exact zeros, so branching
// diverges in nearly every warp, and a zero denominator would push the IEEE division onto its slow path).
template <int METRIC>
__device__ __forceinline__ void dist_step(float& acc, float l, float r) {
    if (METRIC == FIR_L2) {
        float d = __fsub_rn(l, r);
        acc = __fadd_rn(acc, __fmul_rn(d, d));
    } else if (METRIC == FIR_CHI2) {
        const float s = __fadd_rn(l, r);
        const bool pos = s > 0.f;
        const float d = __fsub_rn(l, r);
        // masked lanes divide 1/1: a zero numerator or denominator would send the IEEE division down its slow path
        const float q = __fdiv_rn(pos ? __fmul_rn(d, d) : 1.f, pos ? s : 1.f);
        const float nacc = __fadd_rn(acc, q);
        acc = pos ? nacc : acc;
    } else {
        const float s = __fadd_rn(l, r);
        const bool pos = s > 0.f;
        const float ss = pos ? s : 1.f;
        const bool lp = pos && l > 0.f, rp = pos && r > 0.f;
        const float tl = __fmul_rn(l, glibc_logf(__fdiv_rn(__fmul_rn(2.f, lp ? l : 1.f), ss)));
        const float a1 = __fadd_rn(acc, tl);
        acc = lp ? a1 : acc;
        const float tr = __fmul_rn(r, glibc_logf(__fdiv_rn(__fmul_rn(2.f, rp ? r : 1.f), ss)));
        const float a2 = __fadd_rn(acc, tr);
        acc = rp ? a2 : acc;
    }
}

// One KL step for pairs whose query element l is the same in every lane of the warp (db_features.cpp:33-36).
// l == 0 (either sign): l + r is r exactly, the l-term is skipped by the reference's `l > 0` guard and the r-term is
// r·logf(2r / r) = r·logf(2) whenever r > 0 (2r / r is exactly 2 as long as 2r does not overflow — `light_ok` is false for the
// whole warp otherwise) — no division, no logf, and the branch is uniform.  Any other l takes the general step.
// log2c = glibc_logf(2.0f), the reference's own value, computed by the caller with the same code.
__device__ __forceinline__ void kl_step_uniform1(float& acc, float l, float r, float log2c, bool light_ok) {
    if (l == 0.f && light_ok) {
        const float a = __fadd_rn(acc, __fmul_rn(r, log2c));
        acc = r > 0.f ? a : acc;
    } else dist_step<FIR_KL>(acc, l, r);
}
__device__ __forceinline__ void kl_step_uniform(float& acc0, float& acc1, float l, float r0, float r1, float log2c, bool light_ok) {
    if (l == 0.f && light_ok) {
        const float a0 = __fadd_rn(acc0, __fmul_rn(r0, log2c)), a1 = __fadd_rn(acc1, __fmul_rn(r1, log2c));
        acc0 = r0 > 0.f ? a0 : acc0;
        acc1 = r1 > 0.f ? a1 : acc1;
    } else {
        dist_step<FIR_KL>(acc0, l, r0);
        dist_step<FIR_KL>(acc1, l, r1);
    }
}

}  // namespace fir
#endif
