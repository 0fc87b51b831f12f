/* synth_common.h — the counter-based generator behind the synthetic workloads of BASELINE.json:
 * every element is a pure function of (seed, row, column), so any rank can materialise any row range of the SAME gallery
 * on its GPU and the host can regenerate the identical bits without a GPU.  Shared verbatim by the CUDA kernel
 * (synth_kernels.cu) and the host twin (synth_host.c); fast-image-recognition_b200/synth.py restates it in numpy.
 *
 *   Philox4x32-10 (Salmon et al., SC'11), key = (seed, 0x46495253), counter = (col >> 1, row low, row high, 0);
 *   column col takes output words 2*(col&1), 2*(col&1)+1; the four 16-bit halves are summed (Irwin–Hall, n = 4) and
 *   centred/scaled to unit variance:  z = (float)(sum - 131070) * FIR_SYNTH_INV_STD   — integer arithmetic and ONE fp32
 *   multiplication, hence bit-identical on every IEEE machine (no transcendental functions).
 *
 *   gallery label of row i (class-major, equal class blocks): (i * C) / N;   query label: philox(seed_label, j) mod C
 *   centroid[c][col] = z(seed + 3, c, col)
 *   value = fl(centroid[label][col] + fl(sigma * z(seed + role, row, col)));  relu → max(value, 0)
 */
#ifndef FIR_SYNTH_COMMON_H
#define FIR_SYNTH_COMMON_H
#include <stdint.h>

#ifdef __CUDACC__
#define FIR_SYNTH_FN __host__ __device__ static inline
#else
#define FIR_SYNTH_FN static inline
#endif

#define FIR_SYNTH_KEY1 0x46495253u
#define FIR_SYNTH_INV_STD 0x1.bb67aep-16f /* fp32 nearest to 1 / sqrt(4 * (65536^2 - 1) / 12) = 2.6428997e-05, bits 0x37ddb3d7 */
enum { FIR_SYNTH_GALLERY = 0, FIR_SYNTH_QUERY = 1, FIR_SYNTH_LABEL = 2, FIR_SYNTH_CENTROID = 3 };

FIR_SYNTH_FN void fir_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* the two unit-variance values of columns (2*pair, 2*pair+1) of `row` */
FIR_SYNTH_FN void fir_synth_z2(uint32_t seed, int64_t row, uint32_t pair, float* z0, float* z1) {
    uint32_t w[4];
    fir_philox4x32_10(pair, (uint32_t)((uint64_t)row & 0xffffffffu), (uint32_t)((uint64_t)row >> 32), 0u, seed, FIR_SYNTH_KEY1, w);
    const int s0 = (int)((w[0] & 0xffffu) + (w[0] >> 16) + (w[1] & 0xffffu) + (w[1] >> 16)) - 131070;
    const int s1 = (int)((w[2] & 0xffffu) + (w[2] >> 16) + (w[3] & 0xffffu) + (w[3] >> 16)) - 131070;
    *z0 = (float)s0 * FIR_SYNTH_INV_STD;
    *z1 = (float)s1 * FIR_SYNTH_INV_STD;
}

FIR_SYNTH_FN int32_t fir_synth_label(uint32_t seed, int role, int64_t row, int64_t n_total, int32_t n_classes) {
    if (role == FIR_SYNTH_GALLERY) return (int32_t)((row * (int64_t)n_classes) / n_total);
    uint32_t w[4];
    fir_philox4x32_10(0u, (uint32_t)((uint64_t)row & 0xffffffffu), (uint32_t)((uint64_t)row >> 32), 1u, seed + FIR_SYNTH_LABEL, FIR_SYNTH_KEY1, w);
    return (int32_t)(w[0] % (uint32_t)n_classes);
}
#endif
