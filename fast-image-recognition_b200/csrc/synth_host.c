/* synth_host.c — host twin of synth_kernels.cu (same synth_common.h arithmetic, bit-identical output), threaded over rows.
 * Used where no GPU may be touched: `bench.py --impl reference` (the reference's CPU arm needs the same gallery in host
 * memory) and the CPU tests that pin the generator.  Built by build.sh into libfir_synth_host.so with plain gcc.
 * This is workload generation, not part of the matching path. */
#include "synth_common.h"
#include <pthread.h>
#include <stdlib.h>

typedef struct {
    float* out; int32_t* labels; int64_t row_lo, r0, r1, n_total; int d, n_classes, role, relu; uint32_t seed; float sigma;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    const int pairs = (j->d + 1) / 2;
    float* cen = (float*)malloc(sizeof(float) * 2 * (size_t)pairs);     /* centroid of the current class (gallery rows are class-major) */
    int32_t cen_class = -1;
    for (int64_t r = j->r0; r < j->r1; ++r) {
        const int64_t row = j->row_lo + r;
        const int32_t c = fir_synth_label(j->seed, j->role, row, j->n_total, j->n_classes);
        if (j->labels) j->labels[r] = c;
        if (c != cen_class) {
            for (int p = 0; p < pairs; ++p) fir_synth_z2(j->seed + (uint32_t)FIR_SYNTH_CENTROID, (int64_t)c, (uint32_t)p, &cen[2 * p], &cen[2 * p + 1]);
            cen_class = c;
        }
        float* o = j->out + r * j->d;
        for (int p = 0; p < pairs; ++p) {
            float z0, z1;
            const float m0 = cen[2 * p], m1 = cen[2 * p + 1];
            fir_synth_z2(j->seed + (uint32_t)j->role, row, (uint32_t)p, &z0, &z1);
            volatile float t0 = j->sigma * z0, t1 = j->sigma * z1;      /* separately rounded product, then sum (no contraction) */
            float v0 = m0 + t0, v1 = m1 + t1;
            if (j->relu) { v0 = v0 > 0.f ? v0 : 0.f; v1 = v1 > 0.f ? v1 : 0.f; }
            o[2 * p] = v0;
            if (2 * p + 1 < j->d) o[2 * p + 1] = v1;
        }
    }
    free(cen);
    return 0;
}

int fir_synth_rows_host(float* out_rows, int32_t* out_labels, int64_t row_lo, int64_t n_rows, int64_t n_total, int32_t d,
                        int32_t n_classes, int32_t role, uint32_t seed, float sigma, int32_t relu, int32_t n_threads) {
    if (!out_rows || n_rows < 0 || d <= 0 || n_classes <= 0 || n_total <= 0 || (role != 0 && role != 1)) return 1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if ((int64_t)n_threads > n_rows) n_threads = n_rows > 0 ? (int)n_rows : 1;
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < n_threads; ++t) {
        job_t j = {out_rows, out_labels, row_lo, n_rows * t / n_threads, n_rows * (t + 1) / n_threads, n_total, d, n_classes, role, relu, seed, sigma};
        jobs[t] = j;
        if (t > 0) pthread_create(&th[t], 0, worker, &jobs[t]);
    }
    worker(&jobs[0]);
    for (int t = 1; t < n_threads; ++t) pthread_join(th[t], 0);
    return 0;
}
