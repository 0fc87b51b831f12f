// synth_kernels.cu — device side of the synthetic-workload generator (synth_common.h): rows [row_lo, row_lo + n_rows) of the
// gallery or query matrix of a BASELINE.json config, written straight into HBM.  One thread = two adjacent columns (one
// Philox call for the noise, one for the class centroid), float2 stores; a warp writes 256 contiguous bytes of a row.
#include "fir_common.cuh"
#include "synth_common.h"

namespace fir {

__global__ void __launch_bounds__(256) synth_rows_kernel(float* __restrict__ out, int32_t* __restrict__ labels, int64_t row_lo, int64_t n_rows,
                                                         int64_t n_total, int d, int ld, int n_classes, int role, uint32_t seed, float sigma, int relu) {
    const int pairs = (d + 1) >> 1;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * pairs) return;
    const int64_t r = t / pairs;
    const uint32_t pair = (uint32_t)(t - r * pairs);
    const int64_t row = row_lo + r;
    const int32_t c = fir_synth_label(seed, role, row, n_total, n_classes);
    if (pair == 0 && labels) labels[r] = c;
    float z0, z1, m0, m1;
    fir_synth_z2(seed + (uint32_t)role, row, pair, &z0, &z1);
    fir_synth_z2(seed + (uint32_t)FIR_SYNTH_CENTROID, (int64_t)c, pair, &m0, &m1);
    float v0 = __fadd_rn(m0, __fmul_rn(sigma, z0)), v1 = __fadd_rn(m1, __fmul_rn(sigma, z1));
    if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
    float* o = out + r * ld + 2 * pair;
    if (2 * (int)pair + 1 < d) {
        if ((ld & 1) == 0) *reinterpret_cast<float2*>(o) = make_float2(v0, v1);
        else { o[0] = v0; o[1] = v1; }
    } else o[0] = v0;
}

}  // namespace fir

extern "C" int fir_synth_rows(float* out_rows, int32_t* out_labels, int64_t row_lo, int64_t n_rows, int64_t n_total, int32_t d,
                              int32_t n_classes, int32_t role, uint32_t seed, float sigma, int32_t relu, void* cuda_stream) {
    using namespace fir;
    if (!out_rows || n_rows < 0 || d <= 0 || n_classes <= 0 || n_total <= 0 || row_lo < 0 || (role != 0 && role != 1))
        return fail(FIR_ERR_BAD_ARG, "fir_synth_rows: bad arguments");
    if (n_rows == 0) return FIR_OK;
    const int64_t threads = n_rows * ((d + 1) / 2);
    synth_rows_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, (cudaStream_t)cuda_stream>>>(out_rows, out_labels, row_lo, n_rows, n_total, d, d, n_classes,
                                                                                              role, seed, sigma, relu);
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}
