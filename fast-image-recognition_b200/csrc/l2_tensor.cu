// l2_tensor.cu — Euclidean brute force as a tensor-core contraction with a certified exact rerank.
//
//   d²(q,x) = ‖q‖² + ‖x‖² − 2·q·xᵀ
//
// Stage 0 (seed pass): the candidate kernel over a 2 % sample of the gallery in a branch-free minima mode gives every
// query a starting threshold for its lists (see SeedPlan below).
// Stage 1 (this file, `l2_candidates_kernel_2cta` / `l2_candidates_kernel`): q·xᵀ on tcgen05 tensor cores — fp16
// operands staged by TMA into 128B-swizzled shared memory, fp32 accumulators in TMEM (two 256-column buffers), one
// elected thread issuing `tcgen05.mma` (cta_group::2: one instruction drives both SMs of a CTA pair).  The epilogue never
// materialises the Q x N matrix: each epilogue thread owns one query row of the accumulator (TMEM lane = row) and one
// column half, reads it with `tcgen05.ld` and keeps that query's R best approximate distances in registers.
// Stage 2 (`tensor_prune_kernel` + exact_kernels.cu: pair_list_kernel): candidates that provably cannot matter are
// dropped; the reference's own fp32 arithmetic (feature_distance, qt_cpp/db_features.cpp:22-42) on the survivors.
// Stage 3 (`tensor_select_kernel`): top-k by (exact distance, index) and a CERTIFICATE that no
// non-candidate can beat the k-th exact distance, from a rigorous bound on |approx − true|
// (Cauchy–Schwarz on the fp16 rounding residuals, which are measured per vector at pack time).
// Queries whose certificate fails take a second, individually seeded tensor pass, and what even that cannot certify is
// re-run through the exact CUDA-core kernel on the GPU, so the returned indices/distances are always bit-identical to
// BruteForce::recognize (qt_cpp/ann.cpp:113-126).
#include "fir_common.cuh"
#include "handles.hpp"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <mutex>

namespace fir {

constexpr int BM = 128;          // queries per CTA tile  (UMMA M)
constexpr int BN = 256;          // gallery rows per tile (UMMA N)
constexpr int BK = 64;           // fp16 elements per 128-byte swizzle row
constexpr int UK = 16;           // UMMA K for 16-bit inputs
constexpr int A_KB_BYTES = BM * BK * 2;   // 16 KiB
constexpr int B_KB_BYTES = BN * BK * 2;   // 32 KiB
constexpr int MAX_RES_KB = 8;    // A stays resident in shared memory when D <= 512
constexpr int TMEM_COLS = 512;   // two 256-column accumulators

bool tensor_path_supported(int d) { return d >= 16; }

// 1 = one CTA per tile (cta_group::1); 2 = CTA pairs (cta_group::2).  FIR_TENSOR_CTAS overrides for A/B measurements.
int tensor_cta_mode() {
    static int mode = [] { const char* e = getenv("FIR_TENSOR_CTAS"); int m = e ? atoi(e) : 2; return (m == 1 || m == 2) ? m : 2; }();
    return mode;
}

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    uint32_t spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if ((++spins & 0xfff) == 0) {          // watchdog: a protocol bug must fault, never hang the GPU
            long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000LL) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): start address >> 4 in [0,14),
// leading byte offset (unused for swizzled K-major, = 1) in [16,30), stride byte offset = 8 rows x 128 B
// = 1024 B (>> 4 = 64) in [32,46), descriptor version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)64 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D=f32 (bits 4-5 = 1), A=B=f16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// ---------------------------------------------------------------------------------------------------
// fp16 shadow copies: h = fp16(x * scale) with |h| < 2^-14 flushed to zero (no reliance on fp16
// subnormal handling), ‖x‖² in fp64 → fp32, and the residual norm ‖x − h/scale‖ that feeds the certificate.
// ---------------------------------------------------------------------------------------------------
// `center` (optional, [d]): the shadow is taken of x − center (a shift common to both sides leaves ‖q − x‖² unchanged and
// shrinks the fp16 rounding residuals when the vectors share a large mean — the pivot-space vectors of directed enumeration);
// `exclude` (optional, [n]): rows that must never become candidates (zero shadow, +inf norm, not counted in the statistics).
__global__ void absmax_kernel(const float* __restrict__ rows, int64_t n, int ld, int d, unsigned int* out_bits, const float* __restrict__ center,
                              const unsigned char* __restrict__ exclude) {
    int64_t total = n * d;
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / d;
        int c = (int)(i - r * d);
        if (exclude && exclude[r]) continue;
        m = fmaxf(m, center ? (float)fabs((double)rows[r * ld + c] - (double)center[c]) : fabsf(rows[r * ld + c]));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}

// meta[0] = absmax bits (in), meta[1] = scale (out, float bits)
__global__ void pack_rows_kernel(const float* __restrict__ rows, int64_t n, int64_t rows_padded, int ld, int d, int dph,
                                 __half* __restrict__ h, float* __restrict__ norm2, float* __restrict__ resid,
                                 unsigned int* meta, unsigned int* stats_bits, int64_t perm_a, int64_t perm_b, float* __restrict__ row_scale,
                                 const float* __restrict__ center, const unsigned char* __restrict__ exclude,
                                 const unsigned int* __restrict__ fold_meta) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows_padded) return;
    // scale: a power of two bringing the largest magnitude into [0.5, 1) — one global value for the gallery (meta[0] holds
    // its max|x|), or one per row for queries (row_scale != nullptr: no separate max pass, each epilogue thread owns a row)
    float amax = 0.f;
    if (row_scale) {
        if (row < n) for (int c = lane; c < d; c += 32) amax = fmaxf(amax, center ? (float)fabs((double)rows[row * ld + c] - (double)center[c]) : fabsf(rows[row * ld + c]));
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    } else amax = __uint_as_float(meta[0]);
    float scale = 1.f;
    if (amax > 0.f && isfinite(amax)) {
        int e;
        frexpf(amax, &e);                 // amax = m * 2^e, m in [0.5,1)  ⇒  amax * 2^-e in [0.5,1)
        scale = ldexpf(1.f, -e);
    }
    if (row_scale && lane == 0) row_scale[row] = scale;
    if (row == 0 && lane == 0) meta[1] = __float_as_uint(scale);
    __half* hr = h + row * dph;
    const int64_t src_row = row < n ? (row * perm_a + perm_b) % n : 0;
    if (row >= n || (exclude && exclude[src_row])) {
        for (int c = lane; c < dph; c += 32) hr[c] = __float2half_rn(0.f);
        if (lane == 0) { norm2[row] = __int_as_float(0x7f800000); resid[row] = 0.f; }
        return;
    }
    const float inv = 1.f / scale;        // exact: power of two
    double n2 = 0.0, r2 = 0.0;
    for (int c = lane; c < dph; c += 32) {
        // the (optionally centred) value in double: its rounding to fp32 and then to fp16 both land in the measured residual
        const double x = c < d ? (center ? (double)rows[src_row * ld + c] - (double)center[c] : (double)rows[src_row * ld + c]) : 0.0;
        float xs = (float)x * scale;
        __half hv = __float2half_rn(xs);
        if (fabsf(__half2float(hv)) < 6.103515625e-05f) hv = __float2half_rn(0.f);
        hr[c] = fold_meta ? __hneg(hv) : hv;         // folded norms: the query side is stored negated (exact), see fold_norm_kernel
        float back = __half2float(hv) * inv;
        double df = x - (double)back;
        n2 += x * x;
        r2 += df * df;
    }
    for (int o = 16; o > 0; o >>= 1) {
        n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        r2 += __shfl_xor_sync(0xffffffffu, r2, o);
    }
    if (fold_meta) {
        // query side of the folded norms: columns d, d + 1 both hold U = scale · 2^(−a−1), so that the contraction adds
        // (T_hi + T_lo)·U = ‖x‖²·s_g·s_q / 2 to −q̂·x̂.  U is a power of two; a row whose U leaves the fp16 normal range cannot be
        // folded: it gets U = 0 and an infinite residual, i.e. an infinite error bound — its certificate fails and it is re-run exactly
        __syncwarp();
        if (lane == 0) {
            const float U = ldexpf(scale, -(int)fold_meta[2] - 1);
            const bool ok = U >= 6.103515625e-05f && U <= 32768.f;
            hr[d] = hr[d + 1] = __float2half_rn(ok ? U : 0.f);
            if (!ok) r2 = __longlong_as_double(0x7ff0000000000000LL);
        }
    }
    if (lane == 0) {
        float nf = __double2float_ru(n2);
        float rf = __double2float_ru(sqrt(r2) * (1.0 + 1e-9));
        norm2[row] = (float)n2;
        resid[row] = rf;
        if (stats_bits) {
            atomicMax(&stats_bits[0], __float_as_uint(__double2float_ru(sqrt((double)nf))));
            atomicMax(&stats_bits[1], __float_as_uint(rf));
        }
    }
}

static size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

size_t tensor_side_bytes(int64_t rows, int d, int row_tile) {
    int64_t rp = ceil_div(rows, row_tile) * row_tile;
    int dph = round_up(d, BK);
    return al256((size_t)rp * dph * 2) + 3 * al256((size_t)rp * 4) + 256 + 1024;
}

// Folded norms.  When the k-blocks of the shadow have at least two spare columns (D mod 64 <= 62: the 32-dimensional pivot
// space of directed enumeration, 200-, 96-dimensional features ...), the row norm rides in them: columns d, d + 1 hold
// T = ‖x‖²·s_g·2^a as an fp16 hi/lo pair (a brings the largest T into [64, 128): hi/lo then carry T to 2⁻²⁰ of that maximum,
// values under 2⁻¹⁴ flushed like everywhere else), the query rows are stored NEGATED with U = s_q·2^(−a−1) in the same two
// columns, and the tensor core delivers  acc'' = ‖x‖²·s_g·s_q/2 − q̂·x̂ = v·(s_g·s_q/2)  directly: the epilogue compares raw
// accumulators (no FFMA, no norm loads — the pass over a K = 32 pivot space is bound by exactly those instructions) and
// multiplies by the power of two 2/(s_g·s_q) only when it flushes a list.  Padding / excluded rows have T = +inf.
__global__ void fold_norm_kernel(__half* __restrict__ h, const float* __restrict__ norm2, int64_t rows_padded, int dph, int d,
                                 unsigned int* __restrict__ meta, const float* __restrict__ stats) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows_padded) return;
    const float sg = __uint_as_float(meta[1]);
    const float tmax = stats[0] * stats[0] * sg;          // stats[0] = max ‖x‖, rounded up
    int e = 0;
    if (tmax > 0.f && isfinite(tmax)) frexpf(tmax, &e);   // tmax = m·2^e, m in [0.5, 1)
    const int a = 7 - e;                                  // largest T in [64, 128)
    if (row == 0) meta[2] = (unsigned int)a;
    const float T = ldexpf(norm2[row] * sg, a);           // exact: powers of two
    __half hi = __float2half_rn(T);
    if (fabsf(__half2float(hi)) < 6.103515625e-05f) hi = __float2half_rn(0.f);
    float lof = isfinite(T) ? T - __half2float(hi) : 0.f;
    __half lo = __float2half_rn(lof);
    if (fabsf(__half2float(lo)) < 6.103515625e-05f) lo = __float2half_rn(0.f);
    h[row * dph + d] = hi;
    h[row * dph + d + 1] = lo;
}

int tensor_fold_norms(TensorSide* side, int d, const float* d_stats, cudaStream_t s) {
    if (side->dph - d < 2 || !d_stats) return FIR_OK;
    fold_norm_kernel<<<(unsigned)ceil_div(side->rows_padded, 256), 256, 0, s>>>(side->h, side->norm2, side->rows_padded, side->dph, d, side->meta, d_stats);
    FIR_CUDA_TRY(cudaGetLastError());
    side->folded = true;
    return FIR_OK;
}

static int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; }

static int tensor_pack_side_ex(const float* rows, int64_t n, int ld, int d, int row_tile, void* buf, TensorSide* out, float* d_stats,
                               bool permute, bool per_row_scale, cudaStream_t s, const float* center, const unsigned char* exclude,
                               const unsigned int* fold_meta);
int tensor_pack_side(const float* rows, int64_t n, int ld, int d, int row_tile, void* buf, TensorSide* out, float* d_stats,
                     bool permute, bool per_row_scale, cudaStream_t s, const float* center, const unsigned char* exclude) {
    return tensor_pack_side_ex(rows, n, ld, d, row_tile, buf, out, d_stats, permute, per_row_scale, s, center, exclude, nullptr);
}
static int tensor_pack_side_ex(const float* rows, int64_t n, int ld, int d, int row_tile, void* buf, TensorSide* out, float* d_stats,
                               bool permute, bool per_row_scale, cudaStream_t s, const float* center, const unsigned char* exclude,
                               const unsigned int* fold_meta) {
    int64_t rp = ceil_div(n, row_tile) * row_tile;
    int dph = round_up(d, BK);
    char* p = (char*)(((uintptr_t)buf + 1023) & ~(uintptr_t)1023);     // TMA global address alignment (>=16B); keep 1 KiB
    out->h = (__half*)p; p += al256((size_t)rp * dph * 2);
    out->norm2 = (float*)p; p += al256((size_t)rp * 4);
    out->resid = (float*)p; p += al256((size_t)rp * 4);
    out->row_scale = (float*)p; p += al256((size_t)rp * 4);
    out->meta = (unsigned int*)p;
    out->rows = n; out->rows_padded = rp; out->dph = dph;
    // Gallery shadow rows are stored in a strided permutation so that rows which are neighbours in the caller's
    // (class-major) order are far apart in scan order: the running top-R threshold then tightens like on i.i.d. data.
    out->perm_a = 1; out->perm_b = 0;
    if (permute && n > 2) {
        int64_t a = (int64_t)((double)n * 0.6180339887498949);
        if (a < 1) a = 1;
        while (gcd64(a, n) != 1) ++a;
        out->perm_a = a % n; out->perm_b = n / 3;
    }
    FIR_CUDA_TRY(cudaMemsetAsync(out->meta, 0, 16, s));
    if (!per_row_scale) {
        int blocks = (int)std::min<int64_t>(1184, ceil_div(n * d, 256 * 8));
        absmax_kernel<<<std::max(blocks, 1), 256, 0, s>>>(rows, n, ld, d, out->meta, center, exclude);
        FIR_CUDA_TRY(cudaGetLastError());
    }
    pack_rows_kernel<<<(unsigned)ceil_div(rp, 8), 256, 0, s>>>(rows, n, rp, ld, d, dph, out->h, out->norm2, out->resid, out->meta,
                                                              (unsigned int*)d_stats, out->perm_a, out->perm_b, per_row_scale ? out->row_scale : nullptr,
                                                              center, exclude, (fold_meta && per_row_scale && dph - d >= 2) ? fold_meta : nullptr);
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}

// ---------------------------------------------------------------------------------------------------
// TMA descriptors (driver entry point fetched through the runtime: no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

int tensor_encode_map(CUtensorMap* map, const __half* base, int64_t rows_padded, int dph, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(FIR_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)dph, (cuuint64_t)rows_padded};
    cuuint64_t strides[1] = {(cuuint64_t)dph * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FIR_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return FIR_OK;
}

// ---------------------------------------------------------------------------------------------------
// Work partition.  Units are CTAs (or CTA pairs); a query block is BM (or 2·BM) queries; an item is (query block, tile);
// a unit's work is a short list of SEGMENTS (query block, tile range).
//  * FULL ROUNDS: while at least `grid` query blocks remain, unit u takes query block r·grid + u and sweeps ALL gallery tiles
//    from tile 0.  Every unit walks the gallery in the same order at the same pace, so a B tile is fetched from HBM once per
//    round and served to the other units from L2 — this is what keeps a 10 GB gallery from being re-streamed per query block.
//  * REMAINDER (the last nqb mod grid query blocks), BALANCED form: cut, in query-block-major item order, into `grid`
//    contiguous, equally sized ranges (perfect balance); a remainder query block is then covered by several consecutive units,
//    one slot each.  Every unit is at a different place of the gallery, which is fine while the shadow fits L2.
//  * REMAINDER, PHASED form (galleries larger than L2, after at least one full round): up to four phases; in a phase each of
//    its query blocks is cut into g equal tile ranges and unit u = j·g + s sweeps range s for the phase's j-th block — the
//    units of a range walk it together, so the gallery is again read from HBM once per phase instead of once per remainder
//    query block (C5: 214 GB of 268 GB per launch were remainder reads).  The phases (g per phase) minimise the total sweep
//    length Σ 1/g over the ways to seat the remaining blocks: 21 blocks on 74 units → 18 blocks × 4 ranges, then
//    3 blocks × 24 ranges = 0.292 sweeps against 0.284 for the perfectly balanced cut.
// ---------------------------------------------------------------------------------------------------
constexpr int kMaxPhases = 4;
constexpr int kMaxRanges = 32;          // ranges per query block in a phase = candidate slots of its queries
struct PartPhase { int qb0, nqb, g; };  // query blocks [qb0, qb0 + nqb) of the remainder, each cut into g tile ranges
struct Partition { int64_t ntiles, nqb, full_rounds, rem_total; int grid; int n_phases; PartPhase ph[kMaxPhases]; };
__host__ __device__ inline int64_t part_rem_lo(const Partition& P, int64_t u) { return u * P.rem_total / P.grid; }
__host__ __device__ inline PartPhase part_phase(const Partition& P, int j) {      // (no dynamic indexing of a kernel parameter)
    PartPhase ph = P.ph[0];
    if (j == 1) ph = P.ph[1];
    if (j == 2) ph = P.ph[2];
    if (j == 3) ph = P.ph[3];
    return ph;
}
// segment i of unit u → true while there is one (lo == hi: this unit sits the segment out)
__host__ __device__ inline bool part_segment(const Partition& P, int64_t u, int64_t i, int64_t& qb, int& lo, int& hi) {
    if (i < P.full_rounds) { qb = i * P.grid + u; lo = 0; hi = (int)P.ntiles; return true; }
    const int64_t j = i - P.full_rounds, base = P.full_rounds * P.grid;
    if (P.n_phases > 0) {
        if (j >= P.n_phases) return false;
        const PartPhase ph = part_phase(P, (int)j);
        qb = -1; lo = hi = 0;
        if (u < (int64_t)ph.nqb * ph.g) {
            const int64_t jq = u / ph.g, r = u - jq * ph.g;
            qb = base + ph.qb0 + jq; lo = (int)(r * P.ntiles / ph.g); hi = (int)((r + 1) * P.ntiles / ph.g);
        }
        return true;
    }
    const int64_t a = part_rem_lo(P, u), b = part_rem_lo(P, u + 1);
    if (a >= b) return false;
    const int64_t q = a / P.ntiles + j;
    const int64_t s0 = a > q * P.ntiles ? a : q * P.ntiles, e0 = b < (q + 1) * P.ntiles ? b : (q + 1) * P.ntiles;
    if (s0 >= e0) return false;
    qb = base + q; lo = (int)(s0 - q * P.ntiles); hi = (int)(e0 - q * P.ntiles);
    return true;
}
// candidate slot unit u writes for query block qb
__host__ __device__ inline int part_slot(const Partition& P, int64_t u, int64_t qb) {
    const int64_t rq = qb - P.full_rounds * P.grid;
    if (rq < 0) return 0;
    if (P.n_phases > 0) {
        for (int j = 0; j < kMaxPhases; ++j) {
            const PartPhase ph = part_phase(P, j);
            if (j < P.n_phases && rq >= ph.qb0 && rq < ph.qb0 + ph.nqb) return (int)(u % ph.g);
        }
        return 0;
    }
    const int64_t first = ((rq * P.ntiles + 1) * P.grid - 1) / P.rem_total;     // first unit whose range reaches this query block
    return (int)(u - first);
}
// the remainder goes through phases when the shadow is larger than this (bytes); smaller ones stay in L2 whatever the order
static int64_t phased_min_bytes() {
    static const int64_t v = [] { const char* e = getenv("FIR_TENSOR_PHASED_MIN_BYTES"); return e ? (int64_t)atoll(e) : ((int64_t)48 << 20); }();
    return v;
}
// seats r query blocks on G units in at most kMaxPhases phases with the least total sweep length Σ 1/g
// (dynamic programme over (blocks left, phases left): a phase with g ranges per block seats min(r, G / g) blocks)
static int phase_plan(int r, int G, int max_ranges, PartPhase* out) {
    if (r <= 0 || r > 1024) return 0;
    std::vector<double> cost((size_t)(r + 1) * (kMaxPhases + 1), 1e30);
    std::vector<int> pick((size_t)(r + 1) * (kMaxPhases + 1), 0);
    auto at = [&](int left, int depth) -> size_t { return (size_t)left * (kMaxPhases + 1) + depth; };
    for (int depth = 0; depth <= kMaxPhases; ++depth) cost[at(0, depth)] = 0.0;
    for (int depth = 1; depth <= kMaxPhases; ++depth)
        for (int left = 1; left <= r; ++left)
            for (int g = std::max(1, std::min(max_ranges, G / left)); g <= std::min(G, max_ranges); ++g) {
                const int seats = std::min(left, G / g);
                if (seats <= 0) break;
                const double c = 1.0 / g + cost[at(left - seats, depth - 1)];
                if (c < cost[at(left, depth)] - 1e-12) { cost[at(left, depth)] = c; pick[at(left, depth)] = g; }
            }
    if (cost[at(r, kMaxPhases)] > 1e29) return 0;
    int n = 0, left = r, depth = kMaxPhases;
    while (left > 0) {
        const int g = pick[at(left, depth)], seats = std::min(left, G / g);
        out[n++] = PartPhase{r - left, seats, g};
        left -= seats; --depth;
    }
    return n;
}
inline Partition make_partition(int64_t nq, int64_t n, int n_sm, int ctas, int64_t row_bytes = 0) {
    Partition P{};
    P.ntiles = ceil_div(n, BN);
    P.nqb = ceil_div(nq, BM * ctas);
    const int64_t units = std::max<int64_t>(1, n_sm / ctas);
    P.full_rounds = P.nqb / units;
    const int64_t rem_qb = P.nqb - P.full_rounds * units;
    P.rem_total = rem_qb * P.ntiles;
    P.grid = (int)(P.full_rounds > 0 ? units : std::min<int64_t>(units, std::max<int64_t>(1, P.rem_total)));
    if (P.full_rounds > 0 && rem_qb > 0 && P.ntiles * BN * row_bytes > phased_min_bytes() && P.ntiles >= 4 * kMaxRanges) {
        // every query's candidate arrays are sized for the largest range count (one slot per range): keep them under ~8 GiB
        // (2 lists per slot, 16 entries of 16 bytes each) — with very many queries the remainder is a negligible part of the work
        const int64_t fit = ((int64_t)8 << 30) / (std::max<int64_t>(nq, 1) * 512);
        const int max_ranges = (int)std::max<int64_t>(4, std::min<int64_t>(kMaxRanges, fit));
        P.n_phases = phase_plan((int)rem_qb, P.grid, max_ranges, P.ph);
    }
    return P;
}

// Walks a unit's items in order with adds and compares only: the MMA issuer is ONE thread, and a 64-bit division per
// item is hundreds of serial instructions — enough to starve the tensor pipe between tiles.  Divisions happen once per
// segment (part_segment).
struct WorkIter {
    const Partition* P;
    int64_t u, qb, seg;
    int tile, lo, hi;
    __device__ __forceinline__ void load_next() {
        bool ok;
        do { ++seg; ok = part_segment(*P, u, seg, qb, lo, hi); } while (ok && lo >= hi);
        if (!ok) qb = -1;
        tile = lo;
    }
    __device__ __forceinline__ void init(const Partition& part, int64_t unit) { P = &part; u = unit; seg = -1; qb = -1; lo = hi = 0; load_next(); }
    __device__ __forceinline__ bool done() const { return qb < 0; }
    __device__ __forceinline__ bool in_full() const { return seg < P->full_rounds; }
    __device__ __forceinline__ int64_t next_qb() const {          // query block of the following item, -1 if none
        if (tile + 1 < hi) return qb;
        int64_t s = seg, q; int a, b; bool ok;
        do { ++s; ok = part_segment(*P, u, s, q, a, b); } while (ok && a >= b);
        return ok ? q : -1;
    }
    __device__ __forceinline__ void advance() { if (++tile >= hi) load_next(); }
};

int tensor_plan(int64_t nq, int64_t n, int n_sm, int ctas, int* grid, int* n_slots, int* min_slots, int64_t row_bytes) {
    const Partition P = make_partition(nq, n, n_sm, ctas, row_bytes);
    int slots = 1, least = P.full_rounds > 0 ? 1 : (1 << 30);
    const int64_t rem_qb = P.nqb - P.full_rounds * P.grid;
    if (P.n_phases > 0) {
        for (int j = 0; j < P.n_phases; ++j) slots = std::max(slots, P.ph[j].g);
    } else {
        for (int64_t rq = 0; rq < rem_qb; ++rq) {
            const int64_t qb = P.full_rounds * P.grid + rq;
            const int64_t last_item = (rq + 1) * P.ntiles - 1;
            const int64_t last_unit = ((last_item + 1) * P.grid - 1) / P.rem_total;
            slots = std::max<int>(slots, part_slot(P, last_unit, qb) + 1);
            least = std::min<int>(least, part_slot(P, last_unit, qb) + 1);
        }
    }
    *grid = P.grid;
    *n_slots = slots;
    if (min_slots) *min_slots = least == (1 << 30) ? 1 : least;
    return FIR_OK;
}

// Host-side self check of the partition (no GPU): every (query block, tile) item is covered exactly once, slots are distinct
// per query block and below n_slots.  0 = consistent; used by tests/test_cabi.py over many shapes.
extern "C" int fir_debug_partition_check(int64_t nq, int64_t n, int n_sm, int ctas, int64_t row_bytes, int* n_phases_out, int* n_slots_out) {
    const Partition P = make_partition(nq, n, n_sm, ctas, row_bytes);
    int grid = 0, n_slots = 0, least = 0;
    tensor_plan(nq, n, n_sm, ctas, &grid, &n_slots, &least, row_bytes);
    if (n_phases_out) *n_phases_out = P.n_phases;
    if (n_slots_out) *n_slots_out = n_slots;
    if (P.nqb * P.ntiles > ((int64_t)1 << 26)) return -1;                       // the check keeps a byte per item
    std::vector<unsigned char> seen((size_t)(P.nqb * P.ntiles), 0);
    std::vector<uint64_t> slot_mask((size_t)P.nqb * 4, 0);                     // 256 slot bits per query block
    for (int64_t u = 0; u < P.grid; ++u) {
        int64_t qb; int lo, hi;
        for (int64_t i = 0; part_segment(P, u, i, qb, lo, hi); ++i) {
            if (lo >= hi) continue;
            if (qb < 0 || qb >= P.nqb || lo < 0 || hi > P.ntiles) return 1;
            for (int t = lo; t < hi; ++t) { unsigned char& c = seen[(size_t)(qb * P.ntiles + t)]; if (c) return 2; c = 1; }
            const int sl = part_slot(P, u, qb);
            if (sl < 0 || sl >= n_slots || sl >= 256) return 3;
            uint64_t& word = slot_mask[(size_t)qb * 4 + (size_t)(sl >> 6)];
            if (word >> (sl & 63) & 1) return 4;                                // two segments of one query block on the same slot
            word |= (uint64_t)1 << (sl & 63);
            if (i > P.full_rounds + kMaxPhases + 2) return 5;
        }
    }
    for (unsigned char c : seen) if (!c) return 6;
    return 0;
}

struct CandParams {
    Partition part;
    int64_t nq, n;
    int64_t perm_a, perm_b;     // shadow row p holds original row (p * perm_a + perm_b) mod n
    int nkb;
    int n_slots;
    const float* gal_norm2;
    const float* qry_norm2;
    const unsigned int* gal_meta;
    const float* qry_row_scale;
    float* cand_val;
    int32_t* cand_idx;
    float* slot_bound;
    const int32_t* skip_if_zero; // optional device counter: nothing to do when it reads 0 (second pass without flagged queries)
#ifdef FIR_MEASURE
    int debug_nolist;           // measurement builds only (-DFIR_MEASURE, FIR_TENSOR_DEBUG_NOLIST=1): thresholds at -inf, nothing is ever
                                // listed — WRONG results; compiled out of the shipped library
#endif
    int mins_only;              // seed pass (R = 4): the four slots are plain minima over the four 32-column groups, no lists
    const float* seed_thr;      // optional [nq]: approximate squared distance above which a row cannot matter (see seed pass)
    // per-class nearest neighbour (CLS = 1, 2): NATURAL-order shadow (class-major rows), see tensor_class_min
    const int32_t* cls_begin;   // [n_classes + 1] row range of every class
    int n_classes;
    unsigned long long* cls_keys;   // [nq][n_classes] packed (ordered v bits, row): pass 1 writes, pass 2 reads
    int32_t* cls_cnt;           // [nq][n_classes] rows within the margin of the class minimum (pass 2)
    int32_t* cls_cand;          // [nq][n_classes][kClsSlots]
    const float* cls_E;         // [nq] approximation error bound of the query
    double cls_rho;             // relative error of the reference's sequential sum
    // full rounds of a gallery larger than L2: the pairs re-align every kSyncTiles tiles (see the TMA producer)
    int nf;                     // folded norms: accumulators ARE v·(s_g·s_q/2) (fold_norm_kernel); CTA-pair kernel only
    int sync_tiles;             // tiles between two re-alignments (a power of two, > kSyncLag)
    unsigned int* sync_ctr;     // [1 + kMaxPhases * kMaxRanges] zeroed before the launch ([0]: full rounds, then one per (phase, range)); nullptr = off
};
constexpr int kSyncTiles = 512;         // default CandParams::sync_tiles: 128 MiB of 512-d shadow between two re-alignments
constexpr int kSyncLag = 4;             // tiles between a pair's arrival and its wait: the counter's round trip hides behind them
constexpr unsigned kSyncTimeoutNs = 400000;   // a pair that is not answered (grid not co-resident) moves on: the barrier can only cost time
constexpr int kClsSlots = 4;

// Running top-R of one query row, UNSORTED, in registers: mx is the current maximum (+inf until the list is full) and
// thr = min(mx, tau) is what a new value has to beat; tau is the query's seed threshold (+inf when there is none).
// A better value replaces one slot holding the maximum, then the maximum is recomputed — 5R mostly independent
// instructions (a sorted insert is a 5R-deep dependent chain, and this code runs with a single warp per scheduler).
template <int R>
__device__ __forceinline__ void topr_replace(float (&lv)[R], int (&li)[R], float& mx, float& thr, float tau, float v, int j) {
    bool placed = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool hit = !placed && (lv[r] == mx);
        lv[r] = hit ? v : lv[r];
        li[r] = hit ? j : li[r];
        placed = placed || hit;
    }
    float m = lv[0];
#pragma unroll
    for (int r = 1; r < R; ++r) m = fmaxf(m, lv[r]);
    mx = m;
    thr = fminf(m, tau);
}

constexpr int EPI_WARPS = 8;                 // 2 per scheduler: warp e owns TMEM lanes 32*(e%4).. and columns 128*(e/4)..
constexpr int EPI_COLS = BN / (EPI_WARPS / 4);   // 128 columns per epilogue warp
constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;

// One accumulator tile (this warp's 32 rows x EPI_COLS columns): v = ‖x‖² − 2·s·acc, keep each row's R smallest.
template <int R, bool NF = false>
__device__ __forceinline__ void epilogue_scan_tile(uint32_t taddr, const float* __restrict__ nxs, int jbase, float negc,
                                                   float (&lv)[R], int (&li)[R], float& mx, float& thr, float tau) {
#pragma unroll 1
    for (int c0 = 0; c0 < EPI_COLS; c0 += 32) {
        uint32_t rr[32];
        tc_ld32(taddr + c0, rr);
        tc_wait_ld();
        float gm[8];                                    // minimum of each group of 4 columns
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            if constexpr (NF) {                         // folded norms: the accumulator is the value (in units of s_g·s_q / 2)
                gm[i >> 2] = fminf(fminf(__uint_as_float(rr[i + 0]), __uint_as_float(rr[i + 1])), fminf(__uint_as_float(rr[i + 2]), __uint_as_float(rr[i + 3])));
            } else {
            const float4 nx4 = *reinterpret_cast<const float4*>(&nxs[c0 + i]);
            // two columns per instruction (FFMA2: fma.rn on both halves, the same bits as two fmaf)
            const float2 v01 = __ffma2_rn(make_float2(negc, negc), make_float2(__uint_as_float(rr[i + 0]), __uint_as_float(rr[i + 1])), make_float2(nx4.x, nx4.y));
            const float2 v23 = __ffma2_rn(make_float2(negc, negc), make_float2(__uint_as_float(rr[i + 2]), __uint_as_float(rr[i + 3])), make_float2(nx4.z, nx4.w));
            const float v0 = v01.x, v1 = v01.y, v2 = v23.x, v3 = v23.y;
            rr[i + 0] = __float_as_uint(v0); rr[i + 1] = __float_as_uint(v1);
            rr[i + 2] = __float_as_uint(v2); rr[i + 3] = __float_as_uint(v3);
            gm[i >> 2] = fminf(fminf(v0, v1), fminf(v2, v3));
            }
        }
        const float vmin = fminf(fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3])), fminf(fminf(gm[4], gm[5]), fminf(gm[6], gm[7])));
        if (__any_sync(0xffffffffu, vmin < thr)) {
            // Rare path.  Which groups of 4 columns hold a hit in ANY lane (8-bit masks OR-reduced across the warp): the group
            // index is then warp-uniform, so fetching its 4 values is a uniform switch and the replacement body exists 4x, not 32x.
            uint32_t hg = 0;
#pragma unroll
            for (int g = 0; g < 8; ++g) hg |= (gm[g] < thr) ? (1u << g) : 0u;
            uint32_t many = __reduce_or_sync(0xffffffffu, hg);
#pragma unroll 1
            while (many) {
                const int g = __ffs(many) - 1;
                many &= many - 1;
                uint32_t b0, b1, b2, b3;
                switch (g) {
#define FIR_CASE(G) case G: b0 = rr[4 * G]; b1 = rr[4 * G + 1]; b2 = rr[4 * G + 2]; b3 = rr[4 * G + 3]; break;
                    FIR_CASE(0) FIR_CASE(1) FIR_CASE(2) FIR_CASE(3) FIR_CASE(4) FIR_CASE(5) FIR_CASE(6)
                    default: b0 = rr[28]; b1 = rr[29]; b2 = rr[30]; b3 = rr[31]; break;
#undef FIR_CASE
                }
                const int j = jbase + c0 + 4 * g;
                if (__uint_as_float(b0) < thr) topr_replace<R>(lv, li, mx, thr, tau, __uint_as_float(b0), j);
                if (__uint_as_float(b1) < thr) topr_replace<R>(lv, li, mx, thr, tau, __uint_as_float(b1), j + 1);
                if (__uint_as_float(b2) < thr) topr_replace<R>(lv, li, mx, thr, tau, __uint_as_float(b2), j + 2);
                if (__uint_as_float(b3) < thr) topr_replace<R>(lv, li, mx, thr, tau, __uint_as_float(b3), j + 3);
            }
        }
    }
}

// Seed pass: no lists, no indices — slot g of the row keeps the minimum over column group g (32 columns) of every tile this
// thread sees.  Branch-free, so a short scan is not dominated by list warm-up the way a top-R scan of a few tiles is.
template <bool NF = false>
__device__ __forceinline__ void epilogue_scan_tile_mins(uint32_t taddr, const float* __restrict__ nxs, float negc, float (&lv)[4]) {
#pragma unroll
    for (int c0 = 0; c0 < EPI_COLS; c0 += 32) {
        uint32_t rr[32];
        tc_ld32(taddr + c0, rr);
        tc_wait_ld();
        float m = lv[c0 >> 5];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            if constexpr (NF) {
                m = fminf(m, fminf(fminf(__uint_as_float(rr[i + 0]), __uint_as_float(rr[i + 1])), fminf(__uint_as_float(rr[i + 2]), __uint_as_float(rr[i + 3]))));
            } else {
            const float4 nx4 = *reinterpret_cast<const float4*>(&nxs[c0 + i]);
            const float2 v01 = __ffma2_rn(make_float2(negc, negc), make_float2(__uint_as_float(rr[i + 0]), __uint_as_float(rr[i + 1])), make_float2(nx4.x, nx4.y));
            const float2 v23 = __ffma2_rn(make_float2(negc, negc), make_float2(__uint_as_float(rr[i + 2]), __uint_as_float(rr[i + 3])), make_float2(nx4.z, nx4.w));
            m = fminf(m, fminf(fminf(v01.x, v01.y), fminf(v23.x, v23.y)));
            }
        }
        lv[c0 >> 5] = m;
    }
}

// write one row's list: cand_val/cand_idx [(qrow * n_slots + slot) * R ..], slot_bound = the list's maximum if it is full
// (vscale: 1, or with folded norms the power of two 2 / (s_g·s_q) that turns a raw accumulator into v — an exact product)
template <int R>
__device__ __forceinline__ void epilogue_flush(const CandParams& p, int64_t qrow, int slot, const float (&lv)[R], const int (&li)[R], float thr, float vscale = 1.f) {
    if (qrow >= p.nq) return;
    const float nqv = p.qry_norm2[qrow];
    const int64_t o = (qrow * p.n_slots + slot) * R;
    thr = __fmul_rn(thr, vscale);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        p.cand_val[o + r] = __fmul_rn(lv[r], vscale) + nqv;
        // shadow position -> original gallery row (the fp16 copy is stored in a strided permutation)
        p.cand_idx[o + r] = li[r] < 0 ? -1 : (int32_t)(((int64_t)li[r] * p.perm_a + p.perm_b) % p.n);
    }
    p.slot_bound[qrow * p.n_slots + slot] = thr + nqv;      // thr = min(list maximum or +inf while not full, seed threshold)
}

template <int R, bool A_RES>
__global__ void __launch_bounds__(NUM_THREADS, 1) l2_candidates_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                               const __grid_constant__ CUtensorMap tmap_b, const CandParams p) {
    constexpr int STAGES = A_RES ? 3 : 4;
    constexpr int STAGE_BYTES = A_RES ? B_KB_BYTES : (A_KB_BYTES + B_KB_BYTES);
    constexpr int A_RES_BYTES = A_RES ? MAX_RES_KB * A_KB_BYTES : 0;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_res = smem;
    unsigned char* stage0 = smem + A_RES_BYTES;
    float* nx_s = reinterpret_cast<float*>(stage0 + STAGES * STAGE_BYTES);          // [2][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(nx_s + 2 * BN);
    uint64_t* full_bar = bars;                    // [STAGES]
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]
    uint64_t* a_full = bars + 2 * STAGES;
    uint64_t* a_empty = a_full + 1;
    uint64_t* tmem_full = a_empty + 1;            // [2]
    uint64_t* tmem_empty = tmem_full + 2;         // [2]
    uint64_t* nx_full = tmem_empty + 2;           // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(nx_full + 2);

    if (p.skip_if_zero && *p.skip_if_zero == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Partition P = p.part;
    const int64_t unit = blockIdx.x;
    WorkIter work0; work0.init(p.part, unit);

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();     // SWIZZLE_128B tiles need 1 KiB alignment
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        mbar_init(smem_u32(a_full), 1);
        mbar_init(smem_u32(a_empty), 1);
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tmem_full[s]), 1); mbar_init(smem_u32(&tmem_empty[s]), EPI_WARPS); mbar_init(smem_u32(&nx_full[s]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        int stage = 0; uint32_t phase = 0; uint32_t a_loads = 0;
        int64_t cur_qb = -1;
        for (WorkIter wi = work0; !wi.done(); wi.advance()) {
            const int64_t qb = wi.qb; const int64_t tile = wi.tile;
            if (A_RES && qb != cur_qb) {
                mbar_wait(smem_u32(a_empty), (a_loads & 1) ^ 1);      // every MMA that read the old A has retired
                mbar_expect_tx(smem_u32(a_full), (uint32_t)p.nkb * A_KB_BYTES);
                for (int kb = 0; kb < p.nkb; ++kb)
                    tma_load_2d(smem_u32(a_res + kb * A_KB_BYTES), &tmap_a, smem_u32(a_full), kb * BK, (int)(qb * BM));
                ++a_loads;
            }
            cur_qb = qb;
            for (int kb = 0; kb < p.nkb; ++kb) {
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                unsigned char* st = stage0 + stage * STAGE_BYTES;
                mbar_expect_tx(smem_u32(&full_bar[stage]), STAGE_BYTES);
                if (!A_RES) tma_load_2d(smem_u32(st + B_KB_BYTES), &tmap_a, smem_u32(&full_bar[stage]), kb * BK, (int)(qb * BM));
                tma_load_2d(smem_u32(st), &tmap_b, smem_u32(&full_bar[stage]), kb * BK, (int)(tile * BN));
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer (single thread) =====
        int stage = 0; uint32_t phase = 0; uint32_t a_uses = 0;
        int as = 0; uint32_t aphase = 0;
        int64_t cur_qb = -1;
        for (WorkIter wi = work0; !wi.done(); wi.advance()) {
            const int64_t qb = wi.qb;
            if (A_RES && qb != cur_qb) { mbar_wait(smem_u32(a_full), a_uses & 1); ++a_uses; }
            cur_qb = qb;
            mbar_wait(smem_u32(&tmem_empty[as]), aphase ^ 1);         // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)as * BN;
            for (int kb = 0; kb < p.nkb; ++kb) {
                mbar_wait(smem_u32(&full_bar[stage]), phase);
                tc_fence_after();
                unsigned char* st = stage0 + stage * STAGE_BYTES;
                const uint64_t bdesc = make_sw128_desc(smem_u32(st));
                const uint64_t adesc = make_sw128_desc(A_RES ? smem_u32(a_res + kb * A_KB_BYTES) : smem_u32(st + B_KB_BYTES));
#pragma unroll
                for (int k = 0; k < BK / UK; ++k)      // +32 bytes (>>4 = 2) per 16-element K step inside the swizzle atom
                    tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc, (kb | k) ? 1u : 0u);
                tc_commit(smem_u32(&empty_bar[stage]));               // frees the smem slot when these MMAs retire
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            tc_commit(smem_u32(&tmem_full[as]));                      // accumulator ready for the epilogue
            if (A_RES) {
                const bool last_of_qb = wi.next_qb() != qb;
                if (last_of_qb) tc_commit(smem_u32(a_empty));
            }
            as ^= 1; if (as == 0) aphase ^= 1;
        }
    } else if (warp == 3) {
        // ===== norm loader: stages each tile's 256 gallery norms into nx_s[as] for the epilogue warps =====
        int as = 0; uint32_t aphase = 0;
        for (WorkIter wi = work0; !wi.done(); wi.advance()) {
            const int64_t tile = wi.tile;
            mbar_wait(smem_u32(&tmem_empty[as]), aphase ^ 1);         // epilogue finished with this buffer pair
            const float4* src = reinterpret_cast<const float4*>(p.gal_norm2 + tile * BN + lane * 8);
            const float4 v0 = src[0], v1 = src[1];
            float4* dst = reinterpret_cast<float4*>(&nx_s[as * BN + lane * 8]);
            dst[0] = v0; dst[1] = v1;
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&nx_full[as]));
            as ^= 1; if (as == 0) aphase ^= 1;
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 warps; warp e = (TMEM lane group e%4, column half e/4); thread = one query row =====
        const int e = warp - 4;
        const int lg = e & 3, half = e >> 2;
        const int row = lg * 32 + lane;
        const float sg = __uint_as_float(p.gal_meta[1]);
        float negc = 0.f;                              // −2 / (gallery scale · this row's query scale), set per query block
        float lv[R]; int li[R]; float thr = __int_as_float(0x7f800000), mx = thr, tau = thr;
        int64_t cur_qb = -1;
        int as = 0; uint32_t aphase = 0;
        for (WorkIter wi = work0; !wi.done(); wi.advance()) {
            const int64_t qb = wi.qb; const int64_t tile = wi.tile;
            if (qb != cur_qb) {
                if (cur_qb >= 0) epilogue_flush<R>(p, (cur_qb * BM + row), part_slot(P, unit, cur_qb) * 2 + half, lv, li, thr);
#pragma unroll
                for (int r = 0; r < R; ++r) { lv[r] = __int_as_float(0x7f800000); li[r] = p.mins_only ? 0 : -1; }
                mx = __int_as_float(0x7f800000);
                cur_qb = qb;
                { const int64_t qr = qb * BM + row; negc = -2.0f / (sg * (qr < p.nq ? p.qry_row_scale[qr] : 1.f));
                  tau = (p.seed_thr && qr < p.nq) ? p.seed_thr[qr] - p.qry_norm2[qr] : mx;
#ifdef FIR_MEASURE
                  if (p.debug_nolist) tau = -mx;
#endif
                  thr = tau; }
            }
            mbar_wait(smem_u32(&nx_full[as]), aphase);
            mbar_wait(smem_u32(&tmem_full[as]), aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + half * EPI_COLS);
            if constexpr (R == 4) {
                if (p.mins_only) epilogue_scan_tile_mins(taddr, nx_s + as * BN + half * EPI_COLS, negc, lv);
                else epilogue_scan_tile<R>(taddr, nx_s + as * BN + half * EPI_COLS, (int)(tile * BN) + half * EPI_COLS, negc, lv, li, mx, thr, tau);
            } else {
                epilogue_scan_tile<R>(taddr, nx_s + as * BN + half * EPI_COLS, (int)(tile * BN) + half * EPI_COLS, negc, lv, li, mx, thr, tau);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(smem_u32(&tmem_empty[as])); }
            as ^= 1; if (as == 0) aphase ^= 1;
        }
        if (cur_qb >= 0) epilogue_flush<R>(p, (cur_qb * BM + row), part_slot(P, unit, cur_qb) * 2 + half, lv, li, thr);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}


// ---- per-class nearest neighbour in the tensor epilogue -------------------------------------------------
// The gallery is class-major, so a thread that walks a query row's accumulator columns in row order sees every class as ONE
// run.  PASS 1 keeps the running minimum of v = ‖x‖² − 2·q̂·x̂ of the current class and, at the class change, folds it into
// cls_keys[q][class] with a 64-bit atomicMin (packed (ordered v, row)).  PASS 2 re-scans with the per-(query, class)
// threshold "approximate class minimum + margin" — every row whose REFERENCE distance could be the class minimum lies under
// it (the prune rule of the top-k path with kth := the class minimum) — and lists those rows in cls_cand[q][class][0..3];
// their exact fp32 distances decide.  A chunk of 32 columns that lies inside the current class takes a branch-free
// minimum; only the chunks that contain a class boundary (warp-uniform test) walk their columns one by one.
struct ClsState { int cls, end; float bv; int bi; float thr; };

__device__ __forceinline__ void cls_flush_min(const CandParams& p, int64_t qrow, ClsState& st) {
    if (st.bi >= 0 && qrow < p.nq)
        atomicMin(&p.cls_keys[qrow * p.n_classes + st.cls], ((unsigned long long)ordered_bits(st.bv) << 32) | (uint32_t)st.bi);
    st.bv = __int_as_float(0x7f800000); st.bi = -1;
}
__device__ __forceinline__ void cls_advance(const CandParams& p, ClsState& st, int row) {      // row >= st.end: the class that holds `row`
    int c = st.cls;
    do { ++c; } while (p.cls_begin[c + 1] <= row);                                              // classes without rows are skipped
    st.cls = c; st.end = p.cls_begin[c + 1];
}
__device__ __forceinline__ float cls_threshold(const CandParams& p, int64_t qrow, int cls) {
    if (qrow >= p.nq) return -__int_as_float(0x7f800000);
    const unsigned long long key = p.cls_keys[qrow * p.n_classes + cls];
    if (key == ~0ull) return -__int_as_float(0x7f800000);
    const double nqv = (double)p.qry_norm2[qrow], E = (double)p.cls_E[qrow];
    const double m = (double)from_ordered_bits((uint32_t)(key >> 32)) + nqv;                    // approximate squared distance of the class minimum
    return __double2float_ru((m + E) * (1.0 + p.cls_rho) / (1.0 - p.cls_rho) + E - nqv);
}
__device__ __forceinline__ void cls_emit(const CandParams& p, int64_t qrow, int cls, int row) {
    if (qrow >= p.nq) return;
    const int64_t cell = qrow * p.n_classes + cls;
    const int slot = atomicAdd(&p.cls_cnt[cell], 1);
    if (slot < kClsSlots) p.cls_cand[cell * kClsSlots + slot] = row;
}

template <int PASS>
__device__ __forceinline__ void epilogue_scan_tile_cls(uint32_t taddr, const float* __restrict__ nxs, int jbase, float negc, const CandParams& p,
                                                       int64_t qrow, ClsState& st) {
#pragma unroll 1
    for (int c0 = 0; c0 < EPI_COLS; c0 += 32) {
        uint32_t rr[32];
        tc_ld32(taddr + c0, rr);
        tc_wait_ld();
        const int j0 = jbase + c0;
        float gm[8];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 nx4 = *reinterpret_cast<const float4*>(&nxs[c0 + i]);
            const float2 v01 = __ffma2_rn(make_float2(negc, negc), make_float2(__uint_as_float(rr[i + 0]), __uint_as_float(rr[i + 1])), make_float2(nx4.x, nx4.y));
            const float2 v23 = __ffma2_rn(make_float2(negc, negc), make_float2(__uint_as_float(rr[i + 2]), __uint_as_float(rr[i + 3])), make_float2(nx4.z, nx4.w));
            const float v0 = v01.x, v1 = v01.y, v2 = v23.x, v3 = v23.y;
            rr[i + 0] = __float_as_uint(v0); rr[i + 1] = __float_as_uint(v1); rr[i + 2] = __float_as_uint(v2); rr[i + 3] = __float_as_uint(v3);
            gm[i >> 2] = fminf(fminf(v0, v1), fminf(v2, v3));
        }
        const float vmin = fminf(fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3])), fminf(fminf(gm[4], gm[5]), fminf(gm[6], gm[7])));
        if (j0 + 32 <= st.end) {                                   // warp-uniform: the whole chunk lies inside the current class
            if (PASS == 1) {
                if (vmin < st.bv) {
                    int at = 31;
#pragma unroll
                    for (int i = 30; i >= 0; --i) at = (__uint_as_float(rr[i]) == vmin) ? i : at;
                    st.bv = vmin; st.bi = j0 + at;
                }
            } else {
                // a lane whose chunk minimum is under its threshold lists that column, blanks it and looks again — almost
                // always once (about 1.1 rows per (query, class) lie inside the margin), so the warp pays one select chain
                // instead of 32 predicated emit bodies
                float cur = vmin;
#pragma unroll 1
                while (__any_sync(0xffffffffu, cur <= st.thr)) {
                    if (cur <= st.thr) {
                        int at = -1;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const bool hit = at < 0 && __uint_as_float(rr[i]) == cur;
                            at = hit ? i : at;
                            rr[i] = hit ? 0x7f800000u : rr[i];
                        }
                        cls_emit(p, qrow, st.cls, j0 + at);
                        float m = __uint_as_float(rr[0]);
#pragma unroll
                        for (int i = 1; i < 32; ++i) m = fminf(m, __uint_as_float(rr[i]));
                        cur = m;
                    }
                }
            }
        } else {                                                   // a class boundary (or the end of the gallery) inside the chunk
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int row = j0 + i;
                if (row >= p.n) break;                             // padding rows of the last tile
                if (row >= st.end) {                               // warp-uniform
                    if (PASS == 1) cls_flush_min(p, qrow, st);
                    cls_advance(p, st, row);
                    if (PASS == 2) st.thr = cls_threshold(p, qrow, st.cls);
                }
                const float v = __uint_as_float(rr[i]);
                if (PASS == 1) { if (v < st.bv) { st.bv = v; st.bi = row; } }
                else if (v <= st.thr) cls_emit(p, qrow, st.cls, row);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2).  Two CTAs of a cluster (the two SMs of a TPC) share every gallery tile:
// each loads HALF of the 256-row B tile (128 rows x 64 k, 16 KiB per k-block) and keeps its own 128 query rows
// of a 256-query block resident; one `tcgen05.mma.cta_group::2` (M = 256) issued by the leader CTA consumes both
// halves and writes each CTA's 128 x 256 accumulator into that CTA's TMEM.  Per SM this halves the L2→SMEM bytes
// per flop and doubles the number of k-blocks the same shared memory keeps in flight, which is what the single-CTA
// kernel is bound by (ncu: tensor pipe 24 % active, TMA-wait dominated).
//   barriers in the LEADER only : full[stage] (TMA bytes of both CTAs), a_full, tmem_empty[2] (8 epilogue warps)
//   barriers in BOTH CTAs       : empty[stage], a_empty, tmem_full[2] (multicast tcgen05.commit), nx_full/nx_empty (local)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    // executed by both CTAs; clearing the peer bit of the barrier address makes the bytes count on CTA 0's barrier
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {     // arrive on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t local_bar) {   // arrive on CTA 0's copy of this barrier
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(0));
    // relaxed: the arriving thread publishes no memory; its TMEM reads are already complete (tcgen05.wait::ld)
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// NOTE: waits stay at CTA scope.  A cluster-scope acquire in the try_wait loop makes ptxas emit CCTL.IVALL (an L1
// invalidate) on every spin (ncu: 12 % of all stall samples); nothing waited for here is ordinary memory written by the
// peer CTA — TMA bytes are published by complete_tx, accumulators by tcgen05.commit + tcgen05.fence::after_thread_sync.
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
constexpr uint32_t kIdesc2 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);   // M = 256
constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;   // 16 KiB

template <int R, bool A_RES, int CLS = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
l2_candidates_kernel_2cta(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const CandParams p) {
    constexpr int STAGES = 6;
    constexpr int STAGE_BYTES = A_RES ? B_HALF_BYTES : (A_KB_BYTES + B_HALF_BYTES);
    constexpr int A_RES_BYTES = A_RES ? MAX_RES_KB * A_KB_BYTES : 0;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_res = smem;
    unsigned char* stage0 = smem + A_RES_BYTES;
    float* nx_s = reinterpret_cast<float*>(stage0 + STAGES * STAGE_BYTES);          // [2][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(nx_s + 2 * BN);
    uint64_t* full_bar = bars;                    // [STAGES]  (leader's copy is the live one)
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]
    uint64_t* a_full = bars + 2 * STAGES;
    uint64_t* a_empty = a_full + 1;
    uint64_t* tmem_full = a_empty + 1;            // [2]
    uint64_t* tmem_empty = tmem_full + 2;         // [2]  (leader's copy is the live one)
    uint64_t* nx_full = tmem_empty + 2;           // [2]
    uint64_t* nx_empty = nx_full + 2;             // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(nx_empty + 2);

    if (p.skip_if_zero && *p.skip_if_zero == 0) return;     // same answer in both CTAs of the pair: nobody waits for a peer
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1;
    const Partition P = p.part;
    const int64_t unit = pair;
    WorkIter work0; work0.init(p.part, unit);

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        mbar_init(smem_u32(a_full), 1);
        mbar_init(smem_u32(a_empty), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tmem_full[s]), 1); mbar_init(smem_u32(&tmem_empty[s]), 2 * EPI_WARPS);
            mbar_init(smem_u32(&nx_full[s]), 1); mbar_init(smem_u32(&nx_empty[s]), EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer (both CTAs; every load reports to the leader's barriers) =====
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        int stage = 0; uint32_t phase = 0; uint32_t a_loads = 0;
        unsigned sync_epoch = 0, sync_members = 0; unsigned int* sync_at = nullptr; int64_t sync_seg = -2;
        int64_t cur_qb = -1;
        for (WorkIter wi = work0; !wi.done(); wi.advance()) {
            const int64_t qb = wi.qb; const int64_t tile = wi.tile;
            const int arow = (int)(qb * (2 * BM) + rank * BM);
            // Re-alignment of the pairs that sweep the same tiles in the same order (all pairs in a full round; the pairs of
            // one tile range in a phase of the remainder).  Pairs
            // that drift apart by more than what L2 holds each stream the gallery from HBM on their own (C5 before this:
            // 806 GB per launch against 51 GB for five shared sweeps).  Every kSyncTiles tiles the leader's producer adds
            // one to a global counter and, kSyncLag tiles later, waits until all pairs of that epoch have arrived — the
            // fastest pairs idle for the few microseconds they were ahead.  The wait is bounded, so a grid that is not
            // co-resident loses time, never progress.
            if (p.sync_ctr != nullptr && leader) {
                if (wi.seg != sync_seg) {                          // a new segment: which counter, how many pairs share it
                    sync_seg = wi.seg;
                    if (wi.in_full()) { sync_at = p.sync_ctr; sync_members = (unsigned)P.grid; }             // one monotone counter over all full rounds
                    else {
                        const int j = (int)(wi.seg - P.full_rounds);
                        const PartPhase ph = part_phase(P, j);
                        const bool on = P.n_phases > 0 && (wi.hi - wi.lo) > 2 * p.sync_tiles;
                        sync_at = on ? p.sync_ctr + 1 + j * kMaxRanges + (int)(unit % ph.g) : nullptr;
                        sync_members = (unsigned)ph.nqb; sync_epoch = 0;
                    }
                }
                const int ph = (wi.tile - wi.lo) & (p.sync_tiles - 1);
                if (sync_at != nullptr && ph == 0) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(sync_at) : "memory"); ++sync_epoch; }
                else if (sync_at != nullptr && ph == kSyncLag) {
                    const unsigned target = sync_epoch * sync_members;
                    unsigned seen, t0 = 0, t1;
                    bool timed = false;
                    while (true) {
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(sync_at) : "memory");
                        if ((int)(seen - target) >= 0) break;
                        asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(t1));
                        if (!timed) { t0 = t1; timed = true; }
                        else if (t1 - t0 > kSyncTimeoutNs) break;
                        __nanosleep(100);
                    }
                }
            }
            if (A_RES && qb != cur_qb) {
                mbar_wait(smem_u32(a_empty), (a_loads & 1) ^ 1);
                if (leader) mbar_expect_tx(smem_u32(a_full), 2u * (uint32_t)p.nkb * A_KB_BYTES);
                for (int kb = 0; kb < p.nkb; ++kb)
                    tma_load_2d_2sm(smem_u32(a_res + kb * A_KB_BYTES), &tmap_a, smem_u32(a_full), kb * BK, arow);
                ++a_loads;
            }
            cur_qb = qb;
            for (int kb = 0; kb < p.nkb; ++kb) {
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                unsigned char* st = stage0 + stage * STAGE_BYTES;
                if (leader) mbar_expect_tx(smem_u32(&full_bar[stage]), 2u * STAGE_BYTES);
                if (!A_RES) tma_load_2d_2sm(smem_u32(st + B_HALF_BYTES), &tmap_a, smem_u32(&full_bar[stage]), kb * BK, arow);
                tma_load_2d_2sm(smem_u32(st), &tmap_b, smem_u32(&full_bar[stage]), kb * BK, (int)(tile * BN + rank * (BN / 2)));
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0 && leader) {
        // ===== MMA issuer: one thread of the leader CTA drives both SMs' tensor cores =====
        int stage = 0; uint32_t phase = 0; uint32_t a_uses = 0;
        int as = 0; uint32_t aphase = 0;
        int64_t cur_qb = -1;
        for (WorkIter wi = work0; !wi.done(); wi.advance()) {
            const int64_t qb = wi.qb;
            if (A_RES && qb != cur_qb) { mbar_wait_cluster(smem_u32(a_full), a_uses & 1); ++a_uses; }
            cur_qb = qb;
            mbar_wait_cluster(smem_u32(&tmem_empty[as]), aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)as * BN;
            for (int kb = 0; kb < p.nkb; ++kb) {
                mbar_wait_cluster(smem_u32(&full_bar[stage]), phase);
                tc_fence_after();
                unsigned char* st = stage0 + stage * STAGE_BYTES;
                const uint64_t bdesc = make_sw128_desc(smem_u32(st));
                const uint64_t adesc = make_sw128_desc(A_RES ? smem_u32(a_res + kb * A_KB_BYTES) : smem_u32(st + B_HALF_BYTES));
#pragma unroll
                for (int k = 0; k < BK / UK; ++k)
                    tc_mma_f16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc2, (kb | k) ? 1u : 0u);
                tc_commit_2sm(smem_u32(&empty_bar[stage]));
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            tc_commit_2sm(smem_u32(&tmem_full[as]));
            if (A_RES) {
                const bool last_of_qb = wi.next_qb() != qb;
                if (last_of_qb) tc_commit_2sm(smem_u32(a_empty));
            }
            as ^= 1; if (as == 0) aphase ^= 1;
        }
    } else if (warp == 3) {
        // ===== norm loader (per CTA) =====
        int as = 0; uint32_t aphase = 0;
        for (WorkIter wi = work0; !wi.done(); wi.advance()) {
            const int64_t tile = wi.tile;
            mbar_wait(smem_u32(&nx_empty[as]), aphase ^ 1);
            const float4* src = reinterpret_cast<const float4*>(p.gal_norm2 + tile * BN + lane * 8);
            const float4 v0 = src[0], v1 = src[1];
            float4* dst = reinterpret_cast<float4*>(&nx_s[as * BN + lane * 8]);
            dst[0] = v0; dst[1] = v1;
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&nx_full[as]));
            as ^= 1; if (as == 0) aphase ^= 1;
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 warps; warp e = (TMEM lane group e%4, column half e/4); thread = one query row =====
        const int e = warp - 4;
        const int lg = e & 3, half = e >> 2;
        const int row = lg * 32 + lane;
        const float sg = __uint_as_float(p.gal_meta[1]);
        float negc = 0.f;                              // −2 / (gallery scale · this row's query scale), set per query block
        float vscale = 1.f;                            // folded norms: 2 / (s_g·s_q), a power of two (raw accumulator → v); else 1
        float lv[R]; int li[R]; float thr = __int_as_float(0x7f800000), mx = thr, tau = thr;
        ClsState cst; cst.cls = -1; cst.end = 0; cst.bv = thr; cst.bi = -1; cst.thr = -thr;
        int64_t cur_qb = -1;
        int as = 0; uint32_t aphase = 0;
        for (WorkIter wi = work0; !wi.done(); wi.advance()) {
            const int64_t qb = wi.qb; const int64_t tile = wi.tile;
            if (qb != cur_qb) {
                if constexpr (CLS != 0) {
                    if (CLS == 1 && cur_qb >= 0) cls_flush_min(p, cur_qb * (2 * BM) + rank * BM + row, cst);
                    cst.cls = -1; cst.end = 0; cst.bv = __int_as_float(0x7f800000); cst.bi = -1; cst.thr = -__int_as_float(0x7f800000);
                } else
                if (cur_qb >= 0) epilogue_flush<R>(p, (cur_qb * (2 * BM) + rank * BM + row), part_slot(P, unit, cur_qb) * 2 + half, lv, li, thr, vscale);
#pragma unroll
                for (int r = 0; r < R; ++r) { lv[r] = __int_as_float(0x7f800000); li[r] = p.mins_only ? 0 : -1; }
                mx = __int_as_float(0x7f800000);
                cur_qb = qb;
                { const int64_t qr = qb * (2 * BM) + rank * BM + row;
                  const float sq = qr < p.nq ? p.qry_row_scale[qr] : 1.f;
                  negc = -2.0f / (sg * sq);
                  tau = (p.seed_thr && qr < p.nq) ? p.seed_thr[qr] - p.qry_norm2[qr] : mx;
                  if (p.nf) { vscale = -negc; tau = __fmul_rn(tau, 0.5f * sg * sq); }        // thresholds live in accumulator units (exact: powers of two)
#ifdef FIR_MEASURE
                  if (p.debug_nolist) tau = -mx;
#endif
                  thr = tau; }
            }
            mbar_wait(smem_u32(&nx_full[as]), aphase);
            mbar_wait_cluster(smem_u32(&tmem_full[as]), aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + half * EPI_COLS);
            if constexpr (CLS != 0) {
                epilogue_scan_tile_cls<CLS>(taddr, nx_s + as * BN + half * EPI_COLS, (int)(tile * BN) + half * EPI_COLS, negc, p,
                                            cur_qb * (2 * BM) + rank * BM + row, cst);
            } else if constexpr (R == 4) {
                if (p.mins_only) { if (p.nf) epilogue_scan_tile_mins<true>(taddr, nullptr, negc, lv); else epilogue_scan_tile_mins<false>(taddr, nx_s + as * BN + half * EPI_COLS, negc, lv); }
                else if (p.nf) epilogue_scan_tile<R, true>(taddr, nullptr, (int)(tile * BN) + half * EPI_COLS, negc, lv, li, mx, thr, tau);
                else epilogue_scan_tile<R>(taddr, nx_s + as * BN + half * EPI_COLS, (int)(tile * BN) + half * EPI_COLS, negc, lv, li, mx, thr, tau);
            } else {
                if (p.nf) epilogue_scan_tile<R, true>(taddr, nullptr, (int)(tile * BN) + half * EPI_COLS, negc, lv, li, mx, thr, tau);
                else epilogue_scan_tile<R>(taddr, nx_s + as * BN + half * EPI_COLS, (int)(tile * BN) + half * EPI_COLS, negc, lv, li, mx, thr, tau);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(smem_u32(&nx_empty[as])); if (leader) mbar_arrive(smem_u32(&tmem_empty[as])); else mbar_arrive_leader(smem_u32(&tmem_empty[as])); }
            as ^= 1; if (as == 0) aphase ^= 1;
        }
        if constexpr (CLS != 0) { if (CLS == 1 && cur_qb >= 0) cls_flush_min(p, cur_qb * (2 * BM) + rank * BM + row, cst); }
        else if (cur_qb >= 0) epilogue_flush<R>(p, (cur_qb * (2 * BM) + rank * BM + row), part_slot(P, unit, cur_qb) * 2 + half, lv, li, thr, vscale);
    }

    tc_fence_before();
    cluster_sync_all();          // neither CTA may exit (or free TMEM) while its peer can still touch its smem / barriers
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

static size_t cand_smem_bytes_2cta(bool a_res) {
    size_t stages = 6 * (size_t)(a_res ? B_HALF_BYTES : (A_KB_BYTES + B_HALF_BYTES));
    return (a_res ? (size_t)MAX_RES_KB * A_KB_BYTES : 0) + stages + 2 * BN * 4 + (2 * 6 + 10) * 8 + 16;
}

static size_t cand_smem_bytes(bool a_res) {
    size_t stages = a_res ? 3 * (size_t)B_KB_BYTES : 4 * (size_t)(A_KB_BYTES + B_KB_BYTES);
    return (a_res ? (size_t)MAX_RES_KB * A_KB_BYTES : 0) + stages + 2 * BN * 4 + 18 * 8 + 16;
}

int launch_tensor_candidates(const TensorSearchArgs& a, cudaStream_t s) {
    CandParams p{};
    p.part = make_partition(a.qry->rows, a.gal->rows, a.n_sm, a.ctas, a.row_bytes > 0 ? a.row_bytes : (int64_t)a.gal->dph * 2);
    p.nq = a.qry->rows; p.n = a.gal->rows;
    p.perm_a = a.gal->perm_a; p.perm_b = a.gal->perm_b;
    p.nkb = a.nkb > 0 ? a.nkb : a.gal->dph / BK;
    p.nf = (a.nf && a.ctas == 2) ? 1 : 0;
    p.n_slots = a.n_slots;
    p.gal_norm2 = a.gal->norm2; p.qry_norm2 = a.qry->norm2;
    p.gal_meta = a.gal->meta; p.qry_row_scale = a.qry->row_scale;
    p.cand_val = a.cand_val; p.cand_idx = a.cand_idx; p.slot_bound = a.slot_bound; p.seed_thr = a.seed_thr; p.skip_if_zero = a.skip_if_zero; p.mins_only = a.mins_only;
    {   // re-alignment of the pairs: only where it pays (the shadow does not fit L2 and there are full rounds) and cannot hurt (whole grid resident)
        static const int sync_on = [] { const char* e = getenv("FIR_TENSOR_SYNC"); return e ? atoi(e) : 1; }();
        static const int sync_tiles = [] { const char* e = getenv("FIR_TENSOR_SYNC_TILES"); const int v = e ? atoi(e) : kSyncTiles;
                                           return (v >= 2 * kSyncLag && (v & (v - 1)) == 0) ? v : kSyncTiles; }();     // (tests shrink it to reach the path at small sizes)
        const int64_t shadow_bytes = (int64_t)a.gal->rows_padded * a.gal->dph * 2;
        p.sync_tiles = sync_tiles;
        p.sync_ctr = (sync_on && a.ctas == 2 && a.sync_ctr && p.part.full_rounds >= 1 && shadow_bytes > phased_min_bytes() && p.part.grid * 2 <= a.n_sm &&
                      p.part.ntiles > 2 * sync_tiles) ? a.sync_ctr : nullptr;
    }
#ifdef FIR_MEASURE
    { static const int nolist = [] { const char* e = getenv("FIR_TENSOR_DEBUG_NOLIST"); return e ? atoi(e) : 0; }(); p.debug_nolist = nolist; }
#endif
    const bool a_res = p.nkb <= MAX_RES_KB;
    const size_t smem = a.ctas == 2 ? cand_smem_bytes_2cta(a_res) : cand_smem_bytes(a_res);
    auto go = [&](auto kern) -> int {
        FIR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<a.grid * a.ctas, NUM_THREADS, smem, s>>>(*a.tmap_a, *a.tmap_b, p);      // a.grid counts work units (CTAs or CTA pairs)
        FIR_CUDA_TRY(cudaGetLastError());
        return FIR_OK;
    };
#define FIR_GO(RR) (a.ctas == 2 ? (a_res ? go(l2_candidates_kernel_2cta<RR, true>) : go(l2_candidates_kernel_2cta<RR, false>)) \
                                : (a_res ? go(l2_candidates_kernel<RR, true>) : go(l2_candidates_kernel<RR, false>)))
    switch (a.R) {
        case 4: return FIR_GO(4);
        case 8: return FIR_GO(8);
        case 16: return FIR_GO(16);
        case 32: return FIR_GO(32);
    }
#undef FIR_GO
    return fail(FIR_ERR_INTERNAL, "unsupported candidate list length");
}

// ---------------------------------------------------------------------------------------------------
// Pruning, selection and certificate.  One warp per query.
//   approx(q,x) = ‖q‖² + ‖x‖² − 2·(q̂·x̂);   true(q,x) = ‖q − x‖²
//   |approx − true| ≤ E = 2(‖δq‖(‖x‖+‖δx‖) + ‖q‖‖δx‖) + 2γ‖q‖‖x‖ + η,   δ = fp16 rounding residual (measured per vector),
//   γ bounds the tensor core's fp32 accumulation error, η the fp32 roundings of the norms and of the fma.
//   The reference distance is fl-sum/D with relative error ≤ ρ = (D+4)·2⁻²⁴ around true/D.
// prune : a candidate whose approx exceeds the k-th smallest approx by more than 2E(1+ρ)-ish is strictly worse than k
//         other candidates in the reference's own arithmetic — it cannot be in the answer, so it is not reranked.
// select: top-k by (exact distance, index) over the reranked survivors.
// certify: every non-candidate row has approx ≥ B = min over lists of the list's maximum, hence
//         reference·D ≥ (B − E)(1 − ρ); certified when that exceeds the k-th exact distance (strictly).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double approx_error_bound(double nq2, double rq, double NX, double RX, int nkb);
__device__ __forceinline__ void err_bounds(const ErrModel& m, int64_t q, double& E, double& rho) {
    rho = m.rel;
    if (m.kind == 0) E = approx_error_bound((double)m.q_norm2[q], (double)m.q_resid[q], (double)m.gal_stats[0], (double)m.gal_stats[1], m.nkb) +
                         m.extra_nx2 * (double)m.gal_stats[0] * (double)m.gal_stats[0];
    else if (m.kind == 2) E = m.abs_coef * ((double)m.q_l1[q] + (double)*m.x_l1_max);
    else if (m.kind == 3) {
        const double mp = fmin((double)m.q_minpos[q], (double)*m.x_minpos);      // +inf when a side has no positive element
        const double lam = fmax(1.0, -log2(fmin(mp, 1.0)));
        E = ((double)m.q_l1[q] + (double)*m.x_l1_max) * (m.abs_coef + m.lam_coef * lam);
    }
    else E = 0.0;
}
__device__ __forceinline__ double approx_error_bound(double nq2, double rq, double NX, double RX, int nkb) {
    const double nqn = sqrt(nq2);
    const double gamma = (double)(4 * nkb + 8) * 2.384185791015625e-07;        // (#UMMA K-steps + 8) · 2⁻²²
    return 2.0 * (rq * (NX + RX) + nqn * RX) + 2.0 * gamma * nqn * NX + 1e-6 * (nq2 + NX * NX + 2.0 * nqn * NX);
}

constexpr int PRUNE_STAGE_MAX = 1024;   // valid candidates per query staged in shared memory (4 warps x 8 KiB)

// How many of a query's candidate lists can hold anything: a query block of a full round is written by ONE unit (2 lists), a
// remainder block by one unit per tile range — but the arrays are sized for the widest block (C5: 2 x 24 lists of 16 for every
// one of the 100k queries, of which 95 % use 2).  prune / select scan only the lists their query's block was given.
struct SlotUse { Partition part; int q_per_block; int lists_per_slot; int on; };
__device__ __forceinline__ int slots_used(const SlotUse& u, int64_t q, int n_slots_total) {
    if (!u.on) return n_slots_total;
    const Partition& P = u.part;
    const int64_t rq = q / u.q_per_block - P.full_rounds * P.grid;
    int slots;
    if (rq < 0) slots = 1;
    else if (P.n_phases > 0) {
        slots = 1;
        for (int j = 0; j < kMaxPhases; ++j) {
            const PartPhase ph = part_phase(P, j);
            if (j < P.n_phases && rq >= ph.qb0 && rq < ph.qb0 + ph.nqb) slots = ph.g;
        }
    } else {
        const int64_t first = ((rq * P.ntiles + 1) * P.grid - 1) / P.rem_total;
        const int64_t last = (((rq + 1) * P.ntiles) * P.grid - 1) / P.rem_total;
        slots = (int)(last - first + 1);
    }
    return min(n_slots_total, slots * u.lists_per_slot);
}

__global__ void __launch_bounds__(128) tensor_prune_kernel(float* __restrict__ cand_val, int32_t* __restrict__ cand_idx, int64_t nq, int rt_in, int k,
                                                           const ErrModel em, const int32_t* __restrict__ n_active,
                                                           uint32_t* __restrict__ pair_cells, int32_t* __restrict__ pair_count, const SlotUse su, int R) {
    extern __shared__ __align__(16) unsigned char prune_smem[];
    __shared__ int s_cnt[4];
    __shared__ int s_off;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * 4 + warp;
    const int rt = rt_in;                                                      // row stride of the candidate arrays
    int base = 0;                                                              // survivors of this warp's query
    if (q < nq && n_active && q >= *n_active) {                                // second pass: rows past the flagged count are padding
        int32_t* ci = cand_idx + q * rt;
        for (int c = lane; c < rt; c += 32) ci[c] = -1;
    } else if (q < nq) {
    float* cv = cand_val + q * rt;
    int32_t* ci = cand_idx + q * rt;
    const int rtu = su.on ? slots_used(su, q, rt / R) * R : rt;               // cells beyond it were never written (idx = -1 from the memset)
    // Stage the VALID candidates compacted into shared memory first: one round trip to global memory instead of one per
    // selection round (a warp serves one query, so nothing else hides that latency when only a few queries are active).
    // (Seeded lists are mostly empty, so the capacity is counted in valid entries; a row with more takes the global path.)
    const int cap = min(rt, PRUNE_STAGE_MAX);
    float* sv = reinterpret_cast<float*>(prune_smem) + (size_t)warp * 2 * cap;
    int32_t* si = reinterpret_cast<int32_t*>(sv + cap);
    int nv = 0;
#pragma unroll 4
    for (int c0 = 0; c0 < rtu; c0 += 32) {
        const int c = c0 + lane;
        const int32_t idx = c < rtu ? ci[c] : -1;
        const float v = c < rtu ? cv[c] : 0.f;
        const uint32_t m = __ballot_sync(0xffffffffu, idx >= 0);
        const int o = nv + __popc(m & ((1u << lane) - 1));
        if (idx >= 0 && o < cap) { sv[o] = v; si[o] = idx; }
        nv += __popc(m);
    }
    __syncwarp();
    const float* rv = sv; const int32_t* ri = si;                              // what the rounds read
    if (nv > cap) { rv = cv; ri = ci; nv = rtu; }
    // k-th smallest approx among valid candidates: k rounds of "smallest value greater than the previous" (duplicates counted)
    float kth = -__int_as_float(0x7f800000);
    int taken = 0;
    while (taken < k) {
        float best = __int_as_float(0x7f800000); int cnt = 0;
        for (int c = lane; c < nv; c += 32) {
            if (ri[c] < 0) continue;
            const float v = rv[c];
            if (v > kth && v < best) best = v;
        }
        for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (!(best < __int_as_float(0x7f800000))) break;                       // fewer than k valid candidates
        for (int c = lane; c < nv; c += 32) cnt += (ri[c] >= 0 && rv[c] == best) ? 1 : 0;
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        kth = best; taken += cnt;
    }
    double E, rho;
    err_bounds(em, q, E, rho);
    // c is dominated when (approx_c − E)(1−ρ) > (kth + E)(1+ρ): then reference(c) > reference(each of the k best-by-approx)
    // (fewer than k valid candidates: keep everything)
    const double cut = taken < k ? __longlong_as_double(0x7ff0000000000000LL) : ((double)kth + E) * (1.0 + rho) / (1.0 - rho) + E;
    // compact the survivors to the front of the query's row (order is irrelevant downstream; tensor_select_kernel and the
    // rerank rely on the survivors being a prefix)
    for (int c0 = 0; c0 < nv; c0 += 32) {
        const int c = c0 + lane;
        int32_t idx = -1; float v = 0.f;
        if (c < nv) { idx = ri[c]; v = rv[c]; }
        const bool keep = idx >= 0 && !((double)v > cut);
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) { const int o = base + __popc(m & ((1u << lane) - 1)); ci[o] = idx; cv[o] = v; }   // o ≤ c: never overwrites an unread slot of a later group
        base += __popc(m);
    }
    __syncwarp();
    for (int c = base + lane; c < rtu; c += 32) ci[c] = -1;
    }
    // the survivors of all queries go on one list so that the rerank's warps are full: cell = q * rt + position.  One
    // atomic per block, not per query: ten thousand same-address atomics serialise in L2 for longer than the kernel's work.
    if (pair_cells) {                                                         // (uniform: every warp of the block gets here)
        if (lane == 0) s_cnt[warp] = base;
        __syncthreads();
        if (threadIdx.x == 0) {
            const int total = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
            s_off = total > 0 ? atomicAdd(pair_count, total) : 0;
        }
        __syncthreads();
        int off = s_off;
        for (int w = 0; w < warp; ++w) off += s_cnt[w];
        for (int c = lane; c < base; c += 32) pair_cells[off + c] = (uint32_t)(q * rt + c);
    }
}

__global__ void __launch_bounds__(128) tensor_select_kernel(const float* __restrict__ cand_exact, const int32_t* __restrict__ cand_idx,
                                                            const float* __restrict__ slot_bound, int64_t nq, int n_slots, int R, int k,
                                                            const ErrModel em, int64_t index_offset, float* __restrict__ out_dist,
                                                            int32_t* __restrict__ out_idx, int32_t* flagged, int32_t* n_flagged, unsigned char* fail_flags,
                                                            float* max_bound, const int32_t* __restrict__ n_active, float* __restrict__ seed_out,
                                                            const SlotUse su) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q >= nq || (n_active && q >= *n_active)) return;
    const int rt = n_slots * R;
    const int lists = su.on ? slots_used(su, q, n_slots) : n_slots;          // lists this query's block was given (the others hold nothing)
    const float* ce = cand_exact + q * rt;
    const int32_t* ci = cand_idx + q * rt;
    float last_d = -1.f; int last_i = -1;
    int found = 0;
    float kth = 0.f;
    for (int r = 0; r < k; ++r) {
        float bd = __int_as_float(0x7f800000); int bi = 0x7fffffff;
        for (int c0 = 0; c0 < rt; c0 += 32) {
            const int c = c0 + lane;
            const int i = c < rt ? ci[c] : -1;
            if (__all_sync(0xffffffffu, i < 0)) break;                         // survivors are a prefix (tensor_prune_kernel)
            if (i < 0) continue;
            const float dd = ce[c];
            if (!(dd < 100000.0f)) continue;                                   // ann.cpp:116: nothing ≥ 100000 is ever accepted
            if (r > 0 && !(dd > last_d || (dd == last_d && i > last_i))) continue;
            if (dd < bd || (dd == bd && i < bi)) { bd = dd; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (bi != 0x7fffffff) {
            if (lane == 0) { out_dist[q * k + r] = bd; out_idx[q * k + r] = (int32_t)(bi + index_offset); }
            last_d = bd; last_i = bi; ++found; kth = bd;
        } else {
            if (lane == 0) { out_dist[q * k + r] = 0.f; out_idx[q * k + r] = -1; }
            last_d = __int_as_float(0x7f800000);
        }
    }
    if (lane != 0) return;
    double B = __longlong_as_double(0x7ff0000000000000LL);
    for (int s = 0; s < lists; ++s) {
        const float b = slot_bound[q * n_slots + s];
        if (b == b && (double)b < B) B = (double)b;                            // NaN = list never written = nothing excluded there
    }
    double E, rho;
    err_bounds(em, q, E, rho);
    bool ok;
    if (isinf(B)) ok = true;                                                   // every row of the gallery was a candidate
    else if (found < k) ok = false;
    else ok = (B - E) * (1.0 - rho) > (double)kth * em.dist_scale * (1.0 + rho);
    if (!ok) {
        int pos = atomicAdd(n_flagged, 1); flagged[pos] = (int32_t)q; if (fail_flags) fail_flags[q] = 1;
        // threshold for the second pass: the k-th best found so far bounds the true k-th best from above, so no row whose
        // approximate value exceeds it by the error margins can matter; with this seed the second pass certifies
        // ((B − E)(1 − ρ) > kth(1 + ρ) holds with B = seed) unless one of its lists overflows
        if (seed_out) seed_out[pos] = found >= k ? __double2float_ru((double)kth * em.dist_scale * (1.0 + rho) / (1.0 - rho) + 2.0 * E)
                                                 : __int_as_float(0x7f800000);
    }
    // (same-address atomics from every query serialise; the bound only ever grows, so most warps can see they have nothing to add)
    if (__float_as_uint((float)E) > __ldcg(reinterpret_cast<const unsigned int*>(max_bound)))
        atomicMax(reinterpret_cast<unsigned int*>(max_bound), __float_as_uint((float)E));
}

static int launch_prune_su(float* cand_val, int32_t* cand_idx, int64_t nq, int rt, int k, const ErrModel& em, cudaStream_t s, const int32_t* n_active,
                           uint32_t* pair_cells, int32_t* pair_count, const SlotUse& su, int R) {
    const size_t smem = (size_t)4 * 2 * std::min(rt, PRUNE_STAGE_MAX) * 4;
    tensor_prune_kernel<<<(unsigned)ceil_div(nq, 4), 128, smem, s>>>(cand_val, cand_idx, nq, rt, k, em, n_active, pair_cells, pair_count, su, R);
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}
int launch_prune(float* cand_val, int32_t* cand_idx, int64_t nq, int rt, int k, const ErrModel& em, cudaStream_t s, const int32_t* n_active,
                 uint32_t* pair_cells, int32_t* pair_count) {
    return launch_prune_su(cand_val, cand_idx, nq, rt, k, em, s, n_active, pair_cells, pair_count, SlotUse{}, 1);
}

static int launch_select_su(const float* cand_exact, const int32_t* cand_idx, const float* slot_bound, int64_t nq, int n_slots, int R, int k,
                            const ErrModel& em, int64_t index_offset, float* out_dist, int32_t* out_idx, int32_t* flagged, int32_t* n_flagged,
                            unsigned char* fail_flags, float* max_bound, cudaStream_t s, const int32_t* n_active, float* seed_out, const SlotUse& su) {
    tensor_select_kernel<<<(unsigned)ceil_div(nq, 4), 128, 0, s>>>(cand_exact, cand_idx, slot_bound, nq, n_slots, R, k, em, index_offset, out_dist,
                                                                   out_idx, flagged, n_flagged, fail_flags, max_bound, n_active, seed_out, su);
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}
int launch_select(const float* cand_exact, const int32_t* cand_idx, const float* slot_bound, int64_t nq, int n_slots, int R, int k,
                  const ErrModel& em, int64_t index_offset, float* out_dist, int32_t* out_idx, int32_t* flagged, int32_t* n_flagged,
                  unsigned char* fail_flags, float* max_bound, cudaStream_t s, const int32_t* n_active, float* seed_out) {
    return launch_select_su(cand_exact, cand_idx, slot_bound, nq, n_slots, R, k, em, index_offset, out_dist, out_idx, flagged, n_flagged, fail_flags,
                            max_bound, s, n_active, seed_out, SlotUse{});
}

// ---------------------------------------------------------------------------------------------------
// Orchestration of one fir_search_topk call on the tensor path
// ---------------------------------------------------------------------------------------------------
static int ensure_gallery_side(fir_gallery* g) {
    if (g->tensor_ready) return FIR_OK;
    size_t bytes = tensor_side_bytes(g->n, g->d, BN);
    // a failed earlier attempt may have left its buffers behind: allocate only what is missing, so nothing leaks on a retry
    if (!g->tensor_buf) FIR_CUDA_TRY(cudaMalloc(&g->tensor_buf, bytes));
    if (!g->d_stats) FIR_CUDA_TRY(cudaMalloc(&g->d_stats, 64));
    FIR_CUDA_TRY(cudaMemsetAsync(g->d_stats, 0, 64, g->stream));
    FIR_TRY(tensor_pack_side(g->rows, g->n, g->dp, g->d, BN, g->tensor_buf, &g->tside, g->d_stats, true, false, g->stream, g->tensor_center, g->tensor_exclude));
    {   // spare columns in the last k-block: the row norms ride in them (fold_norm_kernel); FIR_TENSOR_FOLD=0 keeps the norm loads
        static const int fold_on = [] { const char* e = getenv("FIR_TENSOR_FOLD"); return e ? atoi(e) : 1; }();
        if (fold_on && tensor_cta_mode() == 2) FIR_TRY(tensor_fold_norms(&g->tside, g->d, g->d_stats, g->stream));
    }
    FIR_TRY(tensor_encode_map(&g->tmap_b, g->tside.h, g->tside.rows_padded, g->tside.dph, BN));
    FIR_TRY(tensor_encode_map(&g->tmap_b_half, g->tside.h, g->tside.rows_padded, g->tside.dph, BN / 2));
    g->tensor_ready = true;
    return FIR_OK;
}

// second-chance plumbing ---------------------------------------------------------------------------
__global__ void gather_flagged_rows_kernel(const float* __restrict__ q, int ld, const int32_t* __restrict__ flagged,
                                           const int32_t* __restrict__ n_flagged, int64_t cap, float* __restrict__ out) {
    const int64_t i = blockIdx.x;                       // compact row
    const int64_t nf = min((int64_t)*n_flagged, cap);
    const float* src = i < nf ? q + (int64_t)flagged[i] * ld : nullptr;
    for (int c = threadIdx.x; c < ld; c += blockDim.x) out[i * ld + c] = src ? src[c] : 0.f;
}
// results of the second pass go back to their queries; what it could not certify either (or what did not fit in the
// second pass) is queued for the exact re-run
__global__ void escalation_scatter_kernel(const int32_t* __restrict__ flagged, const int32_t* __restrict__ n_flagged, int64_t cap,
                                          const unsigned char* __restrict__ fail2, const float* __restrict__ d2, const int32_t* __restrict__ i2,
                                          int k, float* __restrict__ od, int32_t* __restrict__ oi, int32_t* __restrict__ final_list,
                                          int32_t* __restrict__ n_final) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nf = *n_flagged;
    if (i >= nf) return;
    const int32_t q = flagged[i];
    if (i < cap && !fail2[i]) {
        for (int r = 0; r < k; ++r) { od[(int64_t)q * k + r] = d2[i * k + r]; oi[(int64_t)q * k + r] = i2[i * k + r]; }
    } else {
        final_list[atomicAdd(n_final, 1)] = q;
    }
}

// Seed thresholds.  A list that starts empty accepts everything until it is full and then tightens like R/i: the first
// tiles of every query block are a burst of replacements that outruns the tensor pipe (the accumulator is only double
// buffered).  So the candidate kernel is first run over a SAMPLE of the gallery (the first S rows of the shadow, which is
// stored in a pseudo-random strided order), keeping only the minimum of every (work unit, 32-column group) — a branch-free
// scan — and each query's m-th smallest group minimum becomes the threshold all of its lists start from: only about
// m·N/S rows of the whole gallery fall below it, spread evenly over the scan.  The seed needs no
// guarantee — slot_bound reports min(seed, list maximum), so a seed that turns out too small fails the certificate and the
// query takes the second pass like any other uncertified one.
struct SeedPlan { int64_t S; int m, R, n_slots, grid; float* cand_val; int32_t* cand_idx; float* slot_bound; float* seed; };

static bool seed_enabled() {
    static int on = [] { const char* e = getenv("FIR_TENSOR_SEED"); return e ? atoi(e) : 1; }();
    return on != 0;
}

static size_t seed_plan(fir_gallery* g, int64_t nq, int k, int ctas, SeedPlan* sp) {
    *sp = SeedPlan{};
    static const int div = [] { const char* e = getenv("FIR_TENSOR_SEED_DIV"); int v = e ? atoi(e) : 50; return v > 0 ? v : 50; }();
    static const int m_override = [] { const char* e = getenv("FIR_TENSOR_SEED_M"); return e ? atoi(e) : 0; }();
    // 2 % of the gallery in whole tiles, at most 64 tiles: about m·N/S rows of the whole gallery fall under the seeded threshold
    // — a few thousand list insertions per query over tens of thousands of tiles at N = 10M, where a 2 % sample would cost
    // 2 % of the whole search (16 ms of 850 at C5)
    // ... and no more than 64·sqrt(N / 10M · 512 / D) tiles: the sample costs ~S·D per query block whatever the gallery size, the
    // insertions it saves ~N / S, so the best S grows like sqrt(N / D) (7 tiles at 100k x 512, 23 at 1.25M x 512 — a shard of C5
    // over 8 GPUs, where a 64-tile sample was 3 % of the step — 64 at 10M x 512 and at the 1M x 32 pivot gallery of DEM, whose
    // 32-entry lists make insertions dear: 4.07 -> 4.40 ms per search with 21 tiles)
    const int64_t sqrt_tiles = std::max<int64_t>(4, (int64_t)std::ceil(64.0 * std::sqrt((double)g->n / 1e7 * 512.0 / (double)std::max(g->d, 1))));
    const int64_t S = std::min<int64_t>(g->n / div / BN * BN, std::min<int64_t>(64, sqrt_tiles) * BN);
    if (!seed_enabled() || S < 4 * BN || nq < 4 * BM) return 0;
    sp->S = S;
    // m >= k would make a too-small seed impossible, but every step down in m removes list replacements from the first
    // pass (C2, k = 10, kernel ms: m=8 0.86, 6 0.84, 4 0.81, 3 0.79, 2 0.77); at 0.3 k a few queries in 10^4 take the
    // second pass on top of the ones the certificate sends there anyway, at 0.2 k several times more
    sp->m = m_override > 0 ? m_override : std::max(2, (3 * k + 9) / 10);
    sp->R = 4;                                            // four column-group minima per (slot, half): see epilogue_scan_tile_mins
    int least = 1;
    tensor_plan(nq, S, g->n_sm, ctas, &sp->grid, &sp->n_slots, &least, (int64_t)g->tside.dph * 2);
    sp->n_slots *= EPI_WARPS / 4;
    return 2 * al256((size_t)nq * sp->n_slots * sp->R * 4) + al256((size_t)nq * sp->n_slots * 4) + al256((size_t)nq * 4) + 1024;
}

static bool seed_take(fir_gallery* g, int64_t nq, SeedPlan* sp) {
    if (!sp->S) return true;
    const size_t cells = (size_t)nq * sp->n_slots * sp->R;
    sp->cand_val = (float*)g->ws.take(cells * 4);
    sp->cand_idx = (int32_t*)g->ws.take(cells * 4);
    sp->slot_bound = (float*)g->ws.take((size_t)nq * sp->n_slots * 4);
    sp->seed = (float*)g->ws.take((size_t)nq * 4);
    return sp->cand_val && sp->cand_idx && sp->slot_bound && sp->seed;
}

// seed[q] = m-th smallest distinct group minimum of query q's sample pass (+inf when there are fewer groups); warp per query
__global__ void __launch_bounds__(128) seed_select_kernel(const float* __restrict__ cand_val, const int32_t* __restrict__ cand_idx, int64_t nq, int rt,
                                                          int m, float* __restrict__ seed) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q >= nq) return;
    const float* cv = cand_val + q * rt;
    const int32_t* ci = cand_idx + q * rt;
    const float inf = __int_as_float(0x7f800000);
    float last = -inf;
    for (int r = 0; r < m; ++r) {
        float b = inf;
        for (int c = lane; c < rt; c += 32) {
            const float v = cv[c];
            if (ci[c] >= 0 && v > last && v < b) b = v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b = fminf(b, __shfl_xor_sync(0xffffffffu, b, o));
        last = b;
        if (b == inf) break;
    }
    if (lane == 0) seed[q] = last;
}

namespace {
struct PassBuffers { void* qbuf; float* cand_val; int32_t* cand_idx; float* cand_exact; float* slot_bound; uint32_t* pair_cells; int32_t* pair_count;
                     int n_slots, min_lists, grid, R; size_t qside; };

size_t pass_bytes(fir_gallery* g, int64_t nq, int R, int ctas, PassBuffers* pb) {
    int grid = 0, n_slots = 1, least = 1;
    tensor_plan(nq, g->n, g->n_sm, ctas, &grid, &n_slots, &least, (int64_t)g->tside.dph * 2);
    n_slots *= EPI_WARPS / 4;                           // one list per (CTA slot, column half)
    pb->n_slots = n_slots; pb->min_lists = least * (EPI_WARPS / 4); pb->grid = grid; pb->R = R;
    pb->qside = tensor_side_bytes(nq, g->d, BM);
    return pb->qside + 4 * al256((size_t)nq * n_slots * R * 4) + al256((size_t)nq * n_slots * 4) + 4096;
}
bool pass_take(fir_gallery* g, int64_t nq, PassBuffers* pb) {
    const size_t cells = (size_t)nq * pb->n_slots * pb->R;
    pb->qbuf = g->ws.take(pb->qside);
    pb->cand_val = (float*)g->ws.take(cells * 4);
    pb->cand_idx = (int32_t*)g->ws.take(cells * 4);
    pb->cand_exact = (float*)g->ws.take(cells * 4);
    pb->slot_bound = (float*)g->ws.take((size_t)nq * pb->n_slots * 4);
    pb->pair_cells = (uint32_t*)g->ws.take(cells * 4);
    pb->pair_count = (int32_t*)g->ws.take(1024);      // [0] listed pairs; words 32.. : re-alignment counters of the candidate kernel
    return pb->qbuf && pb->cand_val && pb->cand_idx && pb->cand_exact && pb->slot_bound && pb->pair_cells && pb->pair_count;
}
// pack → tcgen05 candidates → exact rerank → select + certificate, for nq device-resident fp32 queries
int run_pass(fir_gallery* g, const float* dq, int64_t nq, int k, int ctas, const PassBuffers& pb, int64_t index_offset, float* od, int32_t* oi,
             int32_t* flagged, int32_t* n_flagged, unsigned char* fail_flags, float* max_bound, int prof_kind, const SeedPlan* sp = nullptr,
             const int32_t* n_rows = nullptr, float* seed_out = nullptr, const float* seed_in = nullptr) {
    TensorSide qs;
    const bool first = prof_kind == FIR_KERNEL_L2_CANDIDATES;
    auto* ev_pack = first ? g->prof_begin(FIR_PHASE_PACK_QUERIES) : nullptr;
    const int d_eff = g->tensor_d_eff > 0 ? g->tensor_d_eff : g->d;      // distances over the first d_eff dimensions (recognize_image_bf's prefix)
    const bool nf = g->tside.folded && d_eff == g->d && ctas == 2;          // (a prefix call leaves the norm columns to the zero padding of its queries)
    FIR_TRY(tensor_pack_side_ex(dq, nq, g->dp, d_eff, BM, pb.qbuf, &qs, nullptr, false, true, g->stream, g->tensor_center, nullptr, nf ? g->tside.meta : nullptr));
    CUtensorMap tmap_a;
    FIR_TRY(tensor_encode_map(&tmap_a, qs.h, qs.rows_padded, qs.dph, BM));
    const int rt = pb.n_slots * pb.R;
    FIR_CUDA_TRY(cudaMemsetAsync(pb.cand_idx, 0xFF, (size_t)nq * rt * 4, g->stream));
    FIR_CUDA_TRY(cudaMemsetAsync(pb.slot_bound, 0xFF, (size_t)nq * pb.n_slots * 4, g->stream));
    g->prof_end(ev_pack);
    TensorSearchArgs a{};
    a.skip_if_zero = n_rows;
    a.seed_thr = seed_in;
    TensorSide gal_side = g->tside;
    if (d_eff < g->d) { gal_side.norm2 = g->prefix_norm2; a.nkb = round_up(d_eff, BK) / BK; }     // prefix: its own row norms, fewer k-blocks of the same shadow
    a.row_bytes = (int64_t)g->tside.dph * 2;
    a.nf = nf ? 1 : 0;
    a.gal = &gal_side; a.qry = &qs; a.tmap_a = &tmap_a; a.tmap_b = ctas == 2 ? &g->tmap_b_half : &g->tmap_b; a.ctas = ctas;
    a.n_sm = g->n_sm; a.d = d_eff; a.R = pb.R; a.n_slots = pb.n_slots; a.cand_val = pb.cand_val; a.cand_idx = pb.cand_idx; a.slot_bound = pb.slot_bound; a.grid = pb.grid;
    if (sp && sp->S) {
        auto* evs = g->prof_begin(FIR_PHASE_SEED);
        TensorSide sample = gal_side;
        sample.rows = sp->S; sample.rows_padded = sp->S;
        const int srt = sp->n_slots * sp->R;
        FIR_CUDA_TRY(cudaMemsetAsync(sp->cand_idx, 0xFF, (size_t)nq * srt * 4, g->stream));
        TensorSearchArgs sa = a;
        sa.gal = &sample; sa.R = sp->R; sa.n_slots = sp->n_slots; sa.cand_val = sp->cand_val; sa.cand_idx = sp->cand_idx;
        sa.slot_bound = sp->slot_bound; sa.grid = sp->grid; sa.seed_thr = nullptr; sa.mins_only = 1;
        FIR_TRY(launch_tensor_candidates(sa, g->stream));
        seed_select_kernel<<<(unsigned)ceil_div(nq, 4), 128, 0, g->stream>>>(sp->cand_val, sp->cand_idx, nq, srt, sp->m, sp->seed);
        FIR_CUDA_TRY(cudaGetLastError());
        g->prof_end(evs);
        g->stats.gpu_launches += 2;
        a.seed_thr = sp->seed;
    }
    if (first) {                                                          // re-alignment counters of the pairs (words 32.. of the count block)
        FIR_CUDA_TRY(cudaMemsetAsync(pb.pair_count + 32, 0, 4 * (1 + kMaxPhases * kMaxRanges), g->stream));
        a.sync_ctr = reinterpret_cast<unsigned int*>(pb.pair_count + 32);
    }
    { auto* ev = g->prof_begin(prof_kind); int st_ = launch_tensor_candidates(a, g->stream); g->prof_end(ev); FIR_TRY(st_); }
    ErrModel em{};
    em.kind = 0; em.d = d_eff; em.nkb = round_up(d_eff, BK) / BK; em.q_norm2 = qs.norm2; em.q_resid = qs.resid; em.gal_stats = g->d_stats;
    em.rel = (double)(d_eff + 4) * 5.9604644775390625e-08; em.dist_scale = (double)d_eff;
    // folded norms: the hi/lo pair carries ‖x‖²·s_g to 2⁻²⁰ of the largest (1e-6), and the two extra products pass through the
    // tensor core's fp32 accumulation like the others ((4·nkb + 8)·2⁻²² of a term of size ≤ max‖x‖² / 2 — 3e-6 at nkb = 1)
    em.extra_nx2 = nf ? 1e-6 + (double)(4 * em.nkb + 8) * 2.384185791015625e-07 : 0.0;
    auto* ev1 = first ? g->prof_begin(FIR_PHASE_PRUNE) : nullptr;
    const bool listed = (uint64_t)nq * (uint64_t)rt < 0xffffffffull;          // cells are 32-bit
    if (listed) FIR_CUDA_TRY(cudaMemsetAsync(pb.pair_count, 0, 4, g->stream));
    SlotUse su{};
    if (first) {        // (the second pass serves a compacted query list: its few query blocks all use every list)
        su.part = make_partition(nq, g->n, g->n_sm, ctas, (int64_t)g->tside.dph * 2);
        su.q_per_block = BM * ctas; su.lists_per_slot = EPI_WARPS / 4; su.on = 1;
    }
    FIR_TRY(launch_prune_su(pb.cand_val, pb.cand_idx, nq, rt, k, em, g->stream, n_rows, listed ? pb.pair_cells : nullptr, pb.pair_count, su, pb.R));
    g->prof_end(ev1);
    auto* ev2 = first ? g->prof_begin(FIR_PHASE_RERANK) : nullptr;
    if (listed) FIR_TRY(launch_pair_list(FIR_L2, dq, g->dp, g->rows, g->dp, d_eff, pb.pair_cells, pb.pair_count, (int64_t)nq * rt, rt, pb.cand_idx, pb.cand_exact, g->stream));
    else FIR_TRY(launch_pair_distances(FIR_L2, dq, nq, g->dp, g->rows, g->dp, g->n, d_eff, pb.cand_idx, rt, 0, pb.cand_exact, g->stream));
    g->prof_end(ev2);
    auto* ev3 = first ? g->prof_begin(FIR_PHASE_SELECT) : nullptr;
    FIR_TRY(launch_select_su(pb.cand_exact, pb.cand_idx, pb.slot_bound, nq, pb.n_slots, pb.R, k, em, index_offset, od, oi, flagged, n_flagged, fail_flags,
                             max_bound, g->stream, n_rows, seed_out, su));
    g->prof_end(ev3);
    g->stats.gpu_launches += 5;   // pack, candidates, prune, rerank, select
    return FIR_OK;
}
}  // namespace

// ‖x[:d_eff]‖² per shadow row (fp64 sum → fp32, like pack_rows_kernel), +inf for padding / excluded rows
__global__ void prefix_norm_kernel(const float* __restrict__ rows, int64_t n, int64_t rows_padded, int ld, int d_eff, int64_t perm_a, int64_t perm_b,
                                   const unsigned char* __restrict__ exclude, float* __restrict__ norm2) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows_padded) return;
    const int64_t src = row < n ? (row * perm_a + perm_b) % n : 0;
    if (row >= n || (exclude && exclude[src])) { if (lane == 0) norm2[row] = __int_as_float(0x7f800000); return; }
    double n2 = 0.0;
    for (int c = lane; c < d_eff; c += 32) { const double x = (double)rows[src * ld + c]; n2 += x * x; }
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    if (lane == 0) norm2[row] = (float)n2;
}

int tensor_search_topk(fir_gallery* g, const float* queries, int64_t nq, int k, int memspace, int32_t* out_idx, float* out_dist, int d_end) {
    g->dbg_cand_idx = nullptr; g->dbg_cand_val = nullptr; g->dbg_cand_exact = nullptr;      // the previous call's lists die with its workspace
    FIR_TRY(ensure_gallery_side(g));
    // Prefix distances (recognize_image_bf(.., max_features), db_features.cpp:319-335): the same shadow, k-blocks up to the prefix,
    // queries packed with zeros beyond it (so the dimensions the last k-block carries past the prefix contribute exactly nothing),
    // prefix row norms; the full-length norm / residual maxima in d_stats stay valid upper bounds for the error model.
    const int d_eff = (d_end > 0 && d_end < g->d) ? d_end : g->d;
    g->tensor_d_eff = d_eff;
    if (d_eff < g->d) {
        if (g->tensor_center) return fail(FIR_ERR_UNSUPPORTED, "prefix distances are not available on a centred shadow");
        if (!g->prefix_norm2) FIR_CUDA_TRY(cudaMalloc(&g->prefix_norm2, sizeof(float) * (size_t)g->tside.rows_padded));
        if (g->prefix_d != d_eff) {
            prefix_norm_kernel<<<(unsigned)ceil_div(g->tside.rows_padded, 8), 256, 0, g->stream>>>(g->rows, g->n, g->tside.rows_padded, g->dp, d_eff, g->tside.perm_a,
                                                                                                     g->tside.perm_b, g->tensor_exclude, g->prefix_norm2);
            FIR_CUDA_TRY(cudaGetLastError());
            g->prefix_d = d_eff;
        }
    }
    const int ctas = tensor_cta_mode();
    // Pass 1: each query gets (slots x 2 column halves) short lists over disjoint, pseudo-randomly interleaved parts of the
    // gallery — cheap to maintain (cost ~ R² per list).  A list that would have needed more than R of the k best makes the
    // certificate fail; those few queries get a second tensor pass with long lists, and only what even that cannot
    // certify (mass exact ties) is re-run through the exact CUDA-core kernel.  All counts stay on the device.
    const int64_t cap2 = std::min<int64_t>(nq, std::max<int64_t>(256, nq / 64));   // queries served by the second pass (the rest overflow to the exact re-run)
    PassBuffers p1{}, p2{};
    pass_bytes(g, nq, 4, ctas, &p1);
    pass_bytes(g, cap2, 4, ctas, &p2);
    // list length from the number of lists L a query is spread over: room for twice its fair share k / L of the k best, plus slack
    auto pick_R = [&](int lists, int extra) { int want = (2 * k + lists - 1) / lists + 2 + extra; return want <= 4 ? 4 : (want <= 8 ? 8 : (want <= 16 ? 16 : 32)); };
    // (sized for the queries spread over the FEWEST lists: a full-round query block has one slot, i.e. two lists)
    static const int r1_override = [] { const char* e = getenv("FIR_TENSOR_R1"); const int v = e ? atoi(e) : 0; return (v == 4 || v == 8 || v == 16 || v == 32) ? v : 0; }();
    // A second pass is one more sweep of the whole gallery for a single query block: where a sweep is long (at least one full
    // round) the lists get 8 entries even for k = 1 — at N = 10M four-entry lists leave 69 of 100k queries uncertified (the
    // 4th best of half the gallery sits within the error margin of the best), whose second pass costs 17 ms of an 800 ms search.
    const bool long_sweeps = ceil_div(nq, (int64_t)BM * ctas) >= std::max(1, g->n_sm / ctas);
    int R1 = std::max(pick_R(p1.min_lists, 0), long_sweeps ? 8 : 4);
    if (r1_override) R1 = std::max(R1, r1_override);                         // tuning knob: any R >= the rule's is valid
    // the second pass starts every list from the query's own bound (k-th best found in pass 1 plus the error margins, see
    // tensor_select_kernel), so its lists only ever hold rows that can matter: k plus the rows within the error margin of the
    // k-th best, spread over all the lists of the query — 8 per list overflows only on mass ties (-> exact re-run)
    const int R2 = std::max(8, pick_R(p2.min_lists, 0));
    const bool second = true;
    SeedPlan sp{};
    const size_t seed_bytes = seed_plan(g, nq, k, ctas, &sp);
    // Exact re-run of what two tensor passes could not certify (mass exact ties; normally nothing): the first kFbFast
    // queries of the device-side list are spread over up to 1024 gallery splits each (one or two 64-row tiles per block, so
    // even a single query is served by the whole machine); a second launch covers any overflow with coarse splits.  Both
    // exit at once when their part of the list is empty.
    const int64_t kFbFast = 64;
    const int nsplit_fast = (int)std::max<int64_t>(1, std::min<int64_t>(1024, ceil_div(g->n, 64)));
    const int nsplit_slow = (int)std::max<int64_t>(1, std::min<int64_t>(16, ceil_div(g->n, 64 * 8)));
    const size_t fb_cells = std::max<size_t>((size_t)kFbFast * nsplit_fast, (size_t)nq * nsplit_slow) * k;
    size_t need = al256(sizeof(float) * (size_t)nq * g->dp) + pass_bytes(g, nq, R1, ctas, &p1) + 4 * al256((size_t)nq * 4) +
                  2 * al256((size_t)nq * k * 4) + 2 * al256(fb_cells * 4) + seed_bytes + 16384;
    if (second) need += pass_bytes(g, cap2, R2, ctas, &p2) + al256(sizeof(float) * (size_t)cap2 * g->dp) + 2 * al256((size_t)cap2 * k * 4) + 2 * al256((size_t)cap2 * 4);
    FIR_TRY(g->ws.reserve(need));
    // fp32 queries, zero padded (for the exact rerank and the certificate fallback)
    const float* dq = nullptr;
    {
        const int d = g->d, dp = g->dp;
        if (memspace == FIR_DEVICE && d == dp) dq = queries;
        else {
            float* buf = (float*)g->ws.take(sizeof(float) * (size_t)nq * dp);
            if (!buf) return fail(FIR_ERR_INTERNAL, "workspace underestimated (tensor queries)");
            if (memspace == FIR_HOST) {
                if (d != dp) FIR_CUDA_TRY(cudaMemsetAsync(buf, 0, sizeof(float) * (size_t)nq * dp, g->stream));
                FIR_CUDA_TRY(cudaMemcpy2DAsync(buf, sizeof(float) * dp, queries, sizeof(float) * d, sizeof(float) * d, (size_t)nq,
                                               cudaMemcpyHostToDevice, g->stream));
            } else {
                FIR_TRY(launch_pad_rows(queries, nq, d, buf, dp, g->stream));
                g->stats.gpu_launches++;
            }
            dq = buf;
        }
    }
    int32_t* flagged = (int32_t*)g->ws.take((size_t)nq * 4);
    float* seed2 = (float*)g->ws.take((size_t)nq * 4);     // by position on the flagged list; NaN = padding row, accepts nothing
    int32_t* final_list = (int32_t*)g->ws.take((size_t)nq * 4);
    float* od = out_dist; int32_t* oi = out_idx;
    if (memspace == FIR_HOST || !out_dist) od = (float*)g->ws.take((size_t)nq * k * 4);
    if (memspace == FIR_HOST) oi = (int32_t*)g->ws.take((size_t)nq * k * 4);
    float* part_d = (float*)g->ws.take(fb_cells * 4);
    int32_t* part_i = (int32_t*)g->ws.take(fb_cells * 4);
    if (!pass_take(g, nq, &p1) || !seed_take(g, nq, &sp) || !seed2 || !flagged || !final_list || !od || !oi || !part_d || !part_i)
        return fail(FIR_ERR_INTERNAL, "workspace underestimated (tensor path)");
    // device counters: [4] flagged after pass 1, [5] max bound, [6] flagged inside pass 2 (unused list), [7] final exact re-runs
    int32_t* n_flagged = reinterpret_cast<int32_t*>(g->d_stats + 4);
    float* max_bound = g->d_stats + 5;
    int32_t* n_flagged2 = reinterpret_cast<int32_t*>(g->d_stats + 6);
    int32_t* n_final = reinterpret_cast<int32_t*>(g->d_stats + 7);
    FIR_CUDA_TRY(cudaMemsetAsync(g->d_stats + 4, 0, 16, g->stream));
    FIR_CUDA_TRY(cudaMemsetAsync(seed2, 0xFF, (size_t)nq * 4, g->stream));

    FIR_TRY(run_pass(g, dq, nq, k, ctas, p1, g->index_offset, od, oi, flagged, n_flagged, nullptr, max_bound, FIR_KERNEL_L2_CANDIDATES, &sp, nullptr, seed2, nullptr));
    const int32_t* exact_list = flagged; const int32_t* exact_count = n_flagged;
    if (second) {
        float* dq2 = (float*)g->ws.take(sizeof(float) * (size_t)cap2 * g->dp);
        float* od2 = (float*)g->ws.take((size_t)cap2 * k * 4);
        int32_t* oi2 = (int32_t*)g->ws.take((size_t)cap2 * k * 4);
        int32_t* flagged2 = (int32_t*)g->ws.take((size_t)cap2 * 4);
        unsigned char* fail2 = (unsigned char*)g->ws.take((size_t)cap2);
        if (!pass_take(g, cap2, &p2) || !dq2 || !od2 || !oi2 || !flagged2 || !fail2) return fail(FIR_ERR_INTERNAL, "workspace underestimated (second pass)");
        auto* evp2 = g->prof_begin(FIR_PHASE_PASS2);
        gather_flagged_rows_kernel<<<(unsigned)cap2, 128, 0, g->stream>>>(dq, g->dp, flagged, n_flagged, cap2, dq2);
        FIR_CUDA_TRY(cudaMemsetAsync(fail2, 0, (size_t)cap2, g->stream));
        FIR_TRY(run_pass(g, dq2, cap2, k, ctas, p2, g->index_offset, od2, oi2, flagged2, n_flagged2, fail2, max_bound, FIR_KERNEL_L2_CANDIDATES_PASS2, nullptr, n_flagged, nullptr, seed2));
        escalation_scatter_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, g->stream>>>(flagged, n_flagged, cap2, fail2, od2, oi2, k, od, oi, final_list, n_final);
        g->prof_end(evp2);
        FIR_CUDA_TRY(cudaGetLastError());
        g->stats.gpu_launches += 2;
        exact_list = final_list; exact_count = n_final;
    }
    // what is still uncertified: exact CUDA-core re-run (device-side count; see above)
    auto* evx = g->prof_begin(FIR_PHASE_EXACT_RERUN);
    FIR_TRY(exact_topk_device(g, dq, nq, k, d_eff, exact_list, exact_count, part_d, part_i, nsplit_fast, od, oi, 0, kFbFast));
    if (nq > kFbFast)
        FIR_TRY(exact_topk_device(g, dq, nq, k, d_eff, exact_list, exact_count, part_d, part_i, nsplit_slow, od, oi, kFbFast, nq - kFbFast));
    g->prof_end(evx);
    g->stats.path_used = FIR_PATH_TENSOR;
    g->stats.n_candidates = p1.n_slots * p1.R;
    g->stats.n_fallback = -1;     // resolved lazily by fir_search_last_stats
    g->dbg_cand_val = p1.cand_val; g->dbg_cand_exact = p1.cand_exact; g->dbg_cand_idx = p1.cand_idx;
    g->dbg_nq = nq; g->dbg_slots = p1.n_slots; g->dbg_R = p1.R; g->dbg_generation = g->ws.generation;
    if (memspace == FIR_HOST) {
        FIR_CUDA_TRY(cudaMemcpyAsync(out_idx, oi, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, g->stream));
        if (out_dist) FIR_CUDA_TRY(cudaMemcpyAsync(out_dist, od, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, g->stream));
        FIR_CUDA_TRY(cudaStreamSynchronize(g->stream));
    }
    return FIR_OK;
}

// ---------------------------------------------------------------------------------------------------
// Per-class nearest neighbour (fir_class_min, Euclidean) on the tensor path — see epilogue_scan_tile_cls.
// ---------------------------------------------------------------------------------------------------
static int ensure_gallery_nat(fir_gallery* g) {
    if (g->nat_ready && g->nat_classes == g->n_classes) return FIR_OK;
    for (size_t i = 1; i < g->h_labels.size(); ++i)
        if (g->h_labels[i] < g->h_labels[i - 1]) return kApproxDeclined;       // not class-major: the exact kernels fold arbitrary label runs
    const int C = g->n_classes;
    std::vector<int32_t> begin((size_t)C + 1, 0);
    for (int32_t l : g->h_labels) begin[(size_t)l + 1]++;
    for (int c = 0; c < C; ++c) begin[(size_t)c + 1] += begin[(size_t)c];
    if (!g->d_stats) { FIR_CUDA_TRY(cudaMalloc(&g->d_stats, 64)); FIR_CUDA_TRY(cudaMemsetAsync(g->d_stats, 0, 64, g->stream)); }
    if (!g->tensor_buf_nat) {
        FIR_CUDA_TRY(cudaMalloc(&g->tensor_buf_nat, tensor_side_bytes(g->n, g->d, BN)));
        FIR_TRY(tensor_pack_side(g->rows, g->n, g->dp, g->d, BN, g->tensor_buf_nat, &g->tside_nat, g->d_stats, /*permute=*/false, false, g->stream));
        FIR_TRY(tensor_encode_map(&g->tmap_nat_half, g->tside_nat.h, g->tside_nat.rows_padded, g->tside_nat.dph, BN / 2));
    }
    if (g->d_cls_begin) { FIR_CUDA_TRY(cudaStreamSynchronize(g->stream)); FIR_CUDA_TRY(cudaFree(g->d_cls_begin)); g->d_cls_begin = nullptr; }
    FIR_CUDA_TRY(cudaMalloc(&g->d_cls_begin, 4 * ((size_t)C + 1)));
    FIR_CUDA_TRY(cudaMemcpyAsync(g->d_cls_begin, begin.data(), 4 * ((size_t)C + 1), cudaMemcpyHostToDevice, g->stream));
    FIR_CUDA_TRY(cudaStreamSynchronize(g->stream));                            // `begin` is a local
    g->nat_classes = C;
    g->nat_ready = true;
    return FIR_OK;
}

__global__ void cls_margin_kernel(const float* __restrict__ q_norm2, const float* __restrict__ q_resid, const float* __restrict__ gal_stats, int nkb,
                                  int64_t nq, float* __restrict__ E) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) E[q] = __double2float_ru(approx_error_bound((double)q_norm2[q], (double)q_resid[q], (double)gal_stats[0], (double)gal_stats[1], nkb));
}

// the listed (cell, slot) entries of all queries on one dense list for the exact rerank (pair_list_kernel): one atomic per block
__global__ void __launch_bounds__(256) cls_compact_kernel(const int32_t* __restrict__ cand, int64_t entries, uint32_t* __restrict__ pair_cells,
                                                          int32_t* __restrict__ pair_count) {
    __shared__ int s_warp[8];
    __shared__ int s_off;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool live = e < entries && cand[e] >= 0;
    const uint32_t m = __ballot_sync(0xffffffffu, live);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) { const int c = s_warp[w]; s_warp[w] = tot; tot += c; }
        s_off = tot > 0 ? atomicAdd(pair_count, tot) : 0;
    }
    __syncthreads();
    if (live) pair_cells[s_off + s_warp[warp] + __popc(m & ((1u << lane) - 1))] = (uint32_t)e;
}

// one thread per (query, class): the exact minimum over the listed rows, strict '<' on (distance, row); a class with more rows
// inside the margin than the list holds goes on the overflow list — its rows are then all evaluated exactly (cls_cell_exact_kernel)
__global__ void cls_pick_kernel(const int32_t* __restrict__ cand, const int32_t* __restrict__ cnt, const float* __restrict__ exact, int64_t cells,
                                int64_t index_offset, float* __restrict__ out_min, int32_t* __restrict__ out_arg,
                                uint32_t* __restrict__ over_list, int32_t* __restrict__ over_count) {
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= cells) return;
    const int c = cnt[cell];
    if (c > kClsSlots) { over_list[atomicAdd(over_count, 1)] = (uint32_t)cell; return; }
    float bd = 100000.0f; int32_t bi = -1;                                    // ann.cpp:116: nothing >= 100000 is ever accepted
    for (int s = 0; s < c; ++s) {
        const int32_t r = cand[cell * kClsSlots + s];
        const float d = exact[cell * kClsSlots + s];
        if (r >= 0 && (d < bd || (d == bd && bi >= 0 && r < bi))) { bd = d; bi = r; }
    }
    out_min[cell] = bd;
    out_arg[cell] = bi < 0 ? -1 : (int32_t)(bi + index_offset);
}

// one warp per overflowed (query, class) cell: the reference's own arithmetic over every row of the class (lane l takes rows
// b + l, b + l + 32, ...; each distance is the sequential fp32 sum), then the (distance, row) minimum of the warp
__global__ void __launch_bounds__(128) cls_cell_exact_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ x, int ldx, int d,
                                                             const int32_t* __restrict__ cls_begin, int n_classes, const uint32_t* __restrict__ over_list,
                                                             const int32_t* __restrict__ over_count, int64_t index_offset,
                                                             float* __restrict__ out_min, int32_t* __restrict__ out_arg) {
    const int lane = threadIdx.x & 31;
    const int total = *over_count;
    for (int w = blockIdx.x * 4 + (threadIdx.x >> 5); w < total; w += gridDim.x * 4) {
        const uint32_t cell = over_list[w];
        const int64_t qi = cell / (uint32_t)n_classes;
        const int c = (int)(cell - (uint32_t)qi * (uint32_t)n_classes);
        const float* qr = q + qi * ldq;
        float bd = 100000.0f; int32_t bi = -1;
        for (int r = cls_begin[c] + lane; r < cls_begin[c + 1]; r += 32) {
            const float* xr = x + (int64_t)r * ldx;
            float acc = 0.f;
            for (int k = 0; k < d; k += 4) {                                   // rows are zero padded to a multiple of 32 floats
                const float4 a = *reinterpret_cast<const float4*>(qr + k), b = *reinterpret_cast<const float4*>(xr + k);
                dist_step<FIR_L2>(acc, a.x, b.x);
                if (k + 1 < d) dist_step<FIR_L2>(acc, a.y, b.y);
                if (k + 2 < d) dist_step<FIR_L2>(acc, a.z, b.z);
                if (k + 3 < d) dist_step<FIR_L2>(acc, a.w, b.w);
            }
            const float dd = __fdiv_rn(acc, (float)d);
            if (dd < bd) { bd = dd; bi = r; }                                  // ascending rows within a lane: strict '<' keeps the lowest
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o); const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || od < bd || (od == bd && oi < bi))) { bd = od; bi = oi; }
        }
        if (lane == 0) { out_min[cell] = bd; out_arg[cell] = bi < 0 ? -1 : (int32_t)(bi + index_offset); }
    }
}

int tensor_class_min(fir_gallery* g, const float* queries, int64_t nq, int memspace, float* out_min, int32_t* out_arg) {
    { const int st = ensure_gallery_nat(g); if (st != FIR_OK) return st; }
    const int C = g->n_classes, d = g->d, dp = g->dp;
    const size_t cells = (size_t)nq * C;
    cudaStream_t s = g->stream;
    const size_t qside = tensor_side_bytes(nq, d, BM);
    if (cells >= 0xffffffffull) return kApproxDeclined;           // cells are addressed with 32 bits
    size_t need = al256(sizeof(float) * (size_t)nq * dp) + qside + al256(4 * (size_t)nq) * 2 + al256(8 * cells) + 2 * al256(4 * cells) + 3 * al256(4 * cells * kClsSlots) +
                  2 * al256(4 * cells) + 65536;
    FIR_TRY(g->ws.reserve(need));
    const float* dq = nullptr;
    if (memspace == FIR_DEVICE && d == dp) dq = queries;
    else {
        float* buf = (float*)g->ws.take(sizeof(float) * (size_t)nq * dp);
        if (!buf) return fail(FIR_ERR_INTERNAL, "workspace underestimated (class-min queries)");
        if (memspace == FIR_HOST) {
            if (d != dp) FIR_CUDA_TRY(cudaMemsetAsync(buf, 0, sizeof(float) * (size_t)nq * dp, s));
            FIR_CUDA_TRY(cudaMemcpy2DAsync(buf, sizeof(float) * dp, queries, sizeof(float) * d, sizeof(float) * d, (size_t)nq, cudaMemcpyHostToDevice, s));
        } else FIR_TRY(launch_pad_rows(queries, nq, d, buf, dp, s));
        dq = buf;
    }
    void* qbuf = g->ws.take(qside);
    float* E = (float*)g->ws.take(4 * (size_t)nq);
    uint32_t* over_list = (uint32_t*)g->ws.take(4 * cells);
    uint32_t* pair_cells = (uint32_t*)g->ws.take(4 * cells * kClsSlots);
    int32_t* fcount = (int32_t*)g->ws.take(256);              // [0] overflowed cells, [1] live (cell, slot) entries
    unsigned long long* keys = (unsigned long long*)g->ws.take(8 * cells);
    int32_t* cnt = (int32_t*)g->ws.take(4 * cells);
    int32_t* cand = (int32_t*)g->ws.take(4 * cells * kClsSlots);
    float* exact = (float*)g->ws.take(4 * cells * kClsSlots);
    float* dmin = out_min; int32_t* darg = out_arg;
    if (memspace == FIR_HOST) { dmin = (float*)g->ws.take(4 * cells); darg = (int32_t*)g->ws.take(4 * cells); }
    if (!qbuf || !E || !over_list || !pair_cells || !fcount || !keys || !cnt || !cand || !exact || !dmin || !darg)
        return fail(FIR_ERR_INTERNAL, "workspace underestimated (class-min tensor path)");
    TensorSide qs;
    FIR_TRY(tensor_pack_side(dq, nq, dp, d, BM, qbuf, &qs, nullptr, false, true, s));
    CUtensorMap tmap_a;
    FIR_TRY(tensor_encode_map(&tmap_a, qs.h, qs.rows_padded, qs.dph, BM));
    CandParams p{};
    p.part = make_partition(nq, g->n, g->n_sm, 2, (int64_t)g->tside_nat.dph * 2);
    p.nq = nq; p.n = g->n; p.perm_a = 1; p.perm_b = 0;
    p.nkb = g->tside_nat.dph / BK; p.n_slots = 1;
    p.gal_norm2 = g->tside_nat.norm2; p.qry_norm2 = qs.norm2; p.gal_meta = g->tside_nat.meta; p.qry_row_scale = qs.row_scale;
    p.cls_begin = g->d_cls_begin; p.n_classes = C; p.cls_keys = keys; p.cls_cnt = cnt; p.cls_cand = cand; p.cls_E = E;
    p.cls_rho = (double)(d + 4) * 5.9604644775390625e-08;
    const bool a_res = p.nkb <= MAX_RES_KB;
    const size_t smem = cand_smem_bytes_2cta(a_res);
    auto go = [&](auto kern) -> int {
        FIR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<p.part.grid * 2, NUM_THREADS, smem, s>>>(tmap_a, g->tmap_nat_half, p);
        FIR_CUDA_TRY(cudaGetLastError());
        return FIR_OK;
    };
    FIR_CUDA_TRY(cudaMemsetAsync(keys, 0xFF, 8 * cells, s));
    FIR_CUDA_TRY(cudaMemsetAsync(cnt, 0, 4 * cells, s));
    FIR_CUDA_TRY(cudaMemsetAsync(cand, 0xFF, 4 * cells * kClsSlots, s));
    FIR_CUDA_TRY(cudaMemsetAsync(fcount, 0, 8, s));
    cls_margin_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, s>>>(qs.norm2, qs.resid, g->d_stats, p.nkb, nq, E);
    { auto* ev = g->prof_begin(FIR_KERNEL_L2_CANDIDATES);
      const int st = a_res ? go(l2_candidates_kernel_2cta<4, true, 1>) : go(l2_candidates_kernel_2cta<4, false, 1>);
      g->prof_end(ev); FIR_TRY(st); }
    { auto* ev = g->prof_begin(FIR_KERNEL_L2_CANDIDATES_PASS2);
      const int st = a_res ? go(l2_candidates_kernel_2cta<4, true, 2>) : go(l2_candidates_kernel_2cta<4, false, 2>);
      g->prof_end(ev); FIR_TRY(st); }
    // the reference's own fp32 arithmetic on the listed rows, then the per-class pick
    { auto* ev = g->prof_begin(FIR_PHASE_RERANK);
      const int64_t entries = (int64_t)cells * kClsSlots;
      int st = FIR_OK;
      if ((uint64_t)entries < 0xffffffffull) {                 // dense list of the live entries, one lane per (query, row) pair
          cls_compact_kernel<<<(unsigned)ceil_div(entries, 256), 256, 0, s>>>(cand, entries, pair_cells, fcount + 1);
          st = launch_pair_list(FIR_L2, dq, dp, g->rows, dp, d, pair_cells, fcount + 1, entries, C * kClsSlots, cand, exact, s);
      } else st = launch_pair_distances(FIR_L2, dq, nq, dp, g->rows, dp, g->n, d, cand, C * kClsSlots, 0, exact, s);
      g->prof_end(ev); FIR_TRY(st); }
    cls_pick_kernel<<<(unsigned)ceil_div((int64_t)cells, 256), 256, 0, s>>>(cand, cnt, exact, (int64_t)cells, g->index_offset, dmin, darg, over_list, fcount);
    // classes with more rows inside the margin than the list holds (about 1e-4 of the cells on clustered data): every row of the
    // class through the reference's arithmetic, one warp per cell, device-side count
    cls_cell_exact_kernel<<<g->n_sm * 8, 128, 0, s>>>(dq, dp, g->rows, dp, d, g->d_cls_begin, C, over_list, fcount, g->index_offset, dmin, darg);
    FIR_CUDA_TRY(cudaGetLastError());
    g->stats.path_used = FIR_PATH_TENSOR;
    g->stats.gpu_launches += 8;
    g->stats.n_fallback = -3;                                 // resolved lazily: fcount lives in the workspace, copied below
    FIR_CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<int32_t*>(g->d_stats + 4), fcount, 4, cudaMemcpyDeviceToDevice, s));
    if (memspace == FIR_HOST) {
        FIR_CUDA_TRY(cudaMemcpyAsync(out_min, dmin, 4 * cells, cudaMemcpyDeviceToHost, s));
        FIR_CUDA_TRY(cudaMemcpyAsync(out_arg, darg, 4 * cells, cudaMemcpyDeviceToHost, s));
        FIR_CUDA_TRY(cudaStreamSynchronize(s));
    }
    return FIR_OK;
}

}  // namespace fir

extern "C" int fir_debug_tensor_candidates(fir_gallery* g, int32_t* n_slots, int32_t* R, int32_t* idx, float* approx, float* exact) {
    using namespace fir;
    if (!g || !g->dbg_cand_idx || g->dbg_generation != g->ws.generation)      // any later call on the handle re-used (or re-allocated) the workspace
        return fail(FIR_ERR_BAD_ARG, "no tensor-path call recorded on this gallery (the lists are valid until the next call on the handle)");
    if (n_slots) *n_slots = g->dbg_slots;
    if (R) *R = g->dbg_R;
    const size_t cells = (size_t)g->dbg_nq * g->dbg_slots * g->dbg_R;
    FIR_CUDA_TRY(cudaStreamSynchronize(g->stream));
    if (idx) FIR_CUDA_TRY(cudaMemcpy(idx, g->dbg_cand_idx, cells * 4, cudaMemcpyDeviceToHost));
    if (approx) FIR_CUDA_TRY(cudaMemcpy(approx, g->dbg_cand_val, cells * 4, cudaMemcpyDeviceToHost));
    if (exact) FIR_CUDA_TRY(cudaMemcpy(exact, g->dbg_cand_exact, cells * 4, cudaMemcpyDeviceToHost));
    return FIR_OK;
}
