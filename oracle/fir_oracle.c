/* TEST INFRASTRUCTURE ONLY — see fir_oracle.h.  Plain-C restatement of the reference matching
 * path; each function cites the reference lines it follows (paths relative to
 * /root/reference/qt_cpp/).  Build: gcc -std=c11 -O2 -ffp-contract=off (no FMA contraction, so the
 * fp32/fp64 operation order below is exactly what executes).
 */
#define _GNU_SOURCE
#include "fir_oracle.h"
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------
 * glibc 2.39 logf.  Algorithm and constants: ARM optimized-routines logf (N=16 table, degree-3
 * polynomial in double), as shipped in glibc >= 2.27 (sysdeps/ieee754/flt-32/e_logf.c,
 * e_logf_data.c); the table below was read back from this image's libm.so.6 (.rodata 0xb7d40).
 * x86_64 glibc selects by ifunc between an SSE2 build (separate mul/add) and a -mfma build in which
 * GCC contracted every a*b+c; both operation orders were taken from the libm disassembly.
 * ------------------------------------------------------------------------------------------ */
static const double LOGF_T[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2},
};
static const double LOGF_LN2 = 0x1.62e42fefa39efp-1;
static const double LOGF_A[3] = {-0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2};

static inline uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

float fir_oracle_logf(float x, int use_fma) {
    uint32_t ix = asuint(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return -INFINITY;
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return NAN;
        ix = asuint(x * 0x1p23f);
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> 19) & 15u);
    int k = (int32_t)tmp >> 23;
    uint32_t iz = ix - (tmp & 0xff800000u);
    double invc = LOGF_T[i][0], logc = LOGF_T[i][1];
    double z = (double)asfloat(iz);
    double r, y0, r2, y;
    if (use_fma) {
        r = fma(z, invc, -1.0);
        y0 = fma((double)k, LOGF_LN2, logc);
        r2 = r * r;
        y = fma(LOGF_A[1], r, LOGF_A[2]);
        y = fma(LOGF_A[0], r2, y);
        y = fma(y, r2, y0 + r);
    } else {
        r = z * invc - 1.0;
        y0 = logc + (double)k * LOGF_LN2;
        r2 = r * r;
        y = LOGF_A[1] * r + LOGF_A[2];
        y = LOGF_A[0] * r2 + y;
        y = y * r2 + (y0 + r);
    }
    return (float)y;
}

static int g_logf_variant = -2;
int fir_oracle_logf_host_variant(void) {
    if (g_logf_variant != -2) return g_logf_variant;
    int ok_f = 1, ok_s = 1;
    uint32_t s = 12345u;
    for (int t = 0; t < 2000000; ++t) {
        s = s * 1664525u + 1013904223u;
        float x = asfloat(0x30000000u + (s >> 4) % 0x10800000u); /* ~[4.6e-10, 4] */
        float h = logf(x);
        if (asuint(h) != asuint(fir_oracle_logf(x, 1))) ok_f = 0;
        if (asuint(h) != asuint(fir_oracle_logf(x, 0))) ok_s = 0;
    }
    /* both builds agree except ~1e-8 of inputs; prefer the FMA build when undecided */
    g_logf_variant = ok_f ? 1 : (ok_s ? 0 : -1);
    return g_logf_variant;
}
static inline float host_logf(float x) {
    int v = fir_oracle_logf_host_variant();
    return fir_oracle_logf(x, v != 0);
}

/* ------------------------------------------------------------------------------------------
 * feature_distance — db_features.cpp:22-42.  fp32, strictly sequential, one final division.
 * ------------------------------------------------------------------------------------------ */
float fir_oracle_distance(int metric, const float* lhs, const float* rhs, int start_pos, int end_pos) {
    float dist = 0;
    if (metric == FIR_ORACLE_L2) {
        for (int i = start_pos; i < end_pos; ++i) dist += (lhs[i] - rhs[i]) * (lhs[i] - rhs[i]);     /* :26 */
    } else if (metric == FIR_ORACLE_CHI2) {
        for (int i = start_pos; i < end_pos; ++i)
            if ((lhs[i] + rhs[i]) > 0)                                                                /* :29 */
                dist += (lhs[i] - rhs[i]) * (lhs[i] - rhs[i]) / (lhs[i] + rhs[i]);                    /* :31 */
    } else {
        for (int i = start_pos; i < end_pos; ++i)
            if ((lhs[i] + rhs[i]) > 0) {                                                              /* :29 */
                if (lhs[i] > 0) dist += lhs[i] * host_logf(2 * lhs[i] / (lhs[i] + rhs[i]));           /* :33-34 */
                if (rhs[i] > 0) dist += rhs[i] * host_logf(2 * rhs[i] / (lhs[i] + rhs[i]));           /* :35-36 */
            }
    }
    dist /= (end_pos - start_pos);                                                                    /* :40 */
    return dist;
}

/* loader normalisation — db_features.cpp:79-101 */
void fir_oracle_normalize_rows(int metric, float* rows, int64_t n, int d) {
    for (int64_t r = 0; r < n; ++r) {
        float* f = rows + r * d;
        float sum = 0;
        for (int i = 0; i < d; ++i) {
            float x = f[i];
            if (fabs((double)fabsf(x)) < 0.0001) x = 0;                                               /* :85-86 */
            f[i] = x;
            if (metric == FIR_ORACLE_L2) sum += x * x;                                                /* :91 */
            else sum += x;                                                                            /* :93 */
        }
        if (metric == FIR_ORACLE_L2) sum = sqrtf(sum);                                                /* :98 */
        for (int i = 0; i < d; ++i) f[i] /= sum;                                                      /* :100-101 */
    }
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int metric, d, k, max_features, tid, nthreads;
    const float *g, *q;
    int64_t n, nq;
    int32_t* out_idx;
    float* out_dist;
} bf_job;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void* bf_worker(void* p) {
    bf_job* J = (bf_job*)p;
    int64_t lo = J->nq * J->tid / J->nthreads, hi = J->nq * (J->tid + 1) / J->nthreads;
    int end = J->max_features > 0 ? J->max_features : J->d;
    for (int64_t i = lo; i < hi; ++i) {
        const float* qv = J->q + i * J->d;
        if (J->k <= 0) { /* argmin: ann.cpp:115-125 / db_features.cpp:322-333 */
            int bestInd = -1;
            double bestDist = 100000;
            for (int64_t j = 0; j < J->n; ++j) {
                double dist = fir_oracle_distance(J->metric, qv, J->g + j * J->d, 0, end);
                if (dist < bestDist) { bestDist = dist; bestInd = (int)j; }
            }
            J->out_idx[i] = bestInd;
            if (J->out_dist) J->out_dist[i] = bestInd < 0 ? 0.f : (float)bestDist;
        } else {         /* k smallest (dist, j), strict '<' keeps the lower index first on ties */
            int k = J->k;
            int32_t* oi = J->out_idx + i * k;
            float* od = J->out_dist + i * k;
            int cnt = 0;
            for (int64_t j = 0; j < J->n; ++j) {
                float dist = fir_oracle_distance(J->metric, qv, J->g + j * J->d, 0, end);
                if (!((double)dist < 100000)) continue;
                if (cnt == k && !(dist < od[k - 1])) continue;
                int pos = cnt < k ? cnt : k - 1;
                while (pos > 0 && dist < od[pos - 1]) { od[pos] = od[pos - 1]; oi[pos] = oi[pos - 1]; --pos; }
                od[pos] = dist; oi[pos] = (int32_t)j;
                if (cnt < k) ++cnt;
            }
            for (int r = cnt; r < k; ++r) { oi[r] = -1; od[r] = 0.f; }
        }
    }
    return 0;
}

static double run_bf(int metric, const float* g, int64_t n, int d, const float* q, int64_t nq, int k,
                     int max_features, int nthreads, int32_t* out_idx, float* out_dist) {
    if (nthreads < 1) nthreads = 1;
    if (metric == FIR_ORACLE_KL) (void)fir_oracle_logf_host_variant(); /* resolve before threads start */
    bf_job* jobs = (bf_job*)calloc((size_t)nthreads, sizeof(bf_job));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    double t0 = now_s();
    for (int t = 0; t < nthreads; ++t) {
        bf_job J = {metric, d, k, max_features, t, nthreads, g, q, n, nq, out_idx, out_dist};
        jobs[t] = J;
        if (nthreads == 1) bf_worker(&jobs[t]);
        else pthread_create(&th[t], 0, bf_worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
    double t1 = now_s();
    free(jobs); free(th);
    return t1 - t0;
}

double fir_oracle_bf(int metric, const float* g, int64_t n, int d, const float* q, int64_t nq,
                     int max_features, int nthreads, int32_t* out_idx, float* out_dist) {
    return run_bf(metric, g, n, d, q, nq, 0, max_features, nthreads, out_idx, out_dist);
}
double fir_oracle_topk(int metric, const float* g, int64_t n, int d, const float* q, int64_t nq, int k,
                       int nthreads, int32_t* out_idx, float* out_dist) {
    return run_bf(metric, g, n, d, q, nq, k, 0, nthreads, out_idx, out_dist);
}

void fir_oracle_class_min(int metric, const float* g, const int32_t* labels, int64_t n, int d, int n_classes,
                          const float* q, int64_t nq, float* out_min, int32_t* out_arg) {
    for (int64_t i = 0; i < nq; ++i) {
        float* om = out_min + i * n_classes;
        int32_t* oa = out_arg + i * n_classes;
        for (int c = 0; c < n_classes; ++c) { om[c] = 100000.f; oa[c] = -1; }
        for (int64_t j = 0; j < n; ++j) {
            float dist = fir_oracle_distance(metric, q + i * d, g + j * d, 0, d);
            int c = labels[j];
            if ((double)dist < (oa[c] < 0 ? 100000.0 : (double)om[c])) { om[c] = dist; oa[c] = (int32_t)j; }
        }
    }
}

void fir_oracle_pnn_div(int metric, const float* g, const int32_t* labels, int64_t n, int d, int n_classes,
                        const float* q, int64_t nq, double var, double* out_scores, int32_t* out_label) {
    for (int64_t i = 0; i < nq; ++i) {
        double* out = out_scores + i * n_classes;
        for (int c = 0; c < n_classes; ++c) out[c] = 0;
        for (int64_t j = 0; j < n; ++j) {
            float dist = fir_oracle_distance(metric, q + i * d, g + j * d, 0, d);
            out[labels[j]] += exp(-(double)dist / (2 * var));
        }
        double mx = -DBL_MAX; int best = -1;
        for (int c = 0; c < n_classes; ++c) {
            out[c] /= (double)n;
            if (mx < out[c]) { mx = out[c]; best = c; }                /* classification.cpp:217-225 */
        }
        out_label[i] = best;
    }
}

/* ------------------------------------------------------------------------------------------
 * kNN / PNN in fp64 — classification.cpp:116-170, 188-226; normalize() :103-105
 * ------------------------------------------------------------------------------------------ */
static double centred_sqdist(const double* x, const double* qv, const double* avg, int d) {
    double dist = 0;
    for (int fi = 0; fi < d; ++fi) {
        double diff = x[fi] - avg[fi];        /* :132 / :205 */
        double val = qv[fi] - avg[fi];        /* :136 / :208 */
        diff -= val;                          /* :137 / :209 */
        dist += diff * diff;                  /* :141 / :211 */
    }
    return dist;
}

typedef struct { double d; int64_t i; } knn_pair;
static int knn_cmp(const void* a, const void* b) {
    const knn_pair *x = (const knn_pair*)a, *y = (const knn_pair*)b;
    if (x->d < y->d) return -1;
    if (y->d < x->d) return 1;
    return (x->i > y->i) - (x->i < y->i);    /* std::sort is unstable (:151); ties are defined by index here */
}

void fir_oracle_knn(const double* train, const int32_t* train_label, int64_t n, int d, int n_classes,
                    const double* avg, const double* q, int64_t nq, int K, int32_t* out_label) {
    knn_pair* pr = (knn_pair*)malloc(sizeof(knn_pair) * (size_t)n);
    float* outputs = (float*)malloc(sizeof(float) * (size_t)n_classes);
    for (int64_t i = 0; i < nq; ++i) {
        for (int64_t t = 0; t < n; ++t) {
            double dist = centred_sqdist(train + t * d, q + i * d, avg, d);
            dist /= d;                                                  /* :143 */
            pr[t].d = dist; pr[t].i = t;
        }
        qsort(pr, (size_t)n, sizeof(knn_pair), knn_cmp);                /* :151-152 */
        for (int c = 0; c < n_classes; ++c) outputs[c] = 0;
        for (int64_t t = 0; t < n; ++t) {                               /* :154-160 */
            int c = train_label[pr[t].i];
            ++outputs[c];
            if (outputs[c] >= K) break;
        }
        float max_output = (float)-DBL_MAX;                             /* :161 (→ -inf as float) */
        int best = -1;
        for (int c = 0; c < n_classes; ++c)
            if (max_output < outputs[c]) { max_output = outputs[c]; best = c; }
        out_label[i] = best;
    }
    free(pr); free(outputs);
}

void fir_oracle_pnn(const double* train, const int32_t* train_label, int64_t n, int d, int n_classes,
                    const double* avg, const double* q, int64_t nq, double* out_scores, int32_t* out_label) {
    double var = 0.00002;                                               /* :190 */
    if (d > 2000) var /= 10;                                            /* :192-193 */
    for (int64_t i = 0; i < nq; ++i) {
        double* outputs = out_scores + i * n_classes;
        for (int c = 0; c < n_classes; ++c) outputs[c] = 0;
        for (int64_t t = 0; t < n; ++t) {                               /* class-major ⇒ same add order as :195-214 */
            double dist = centred_sqdist(train + t * d, q + i * d, avg, d);
            outputs[train_label[t]] += exp(-dist / (2 * (double)(size_t)d * var));   /* :213 */
        }
        double max_output = -DBL_MAX; int best = -1;
        for (int c = 0; c < n_classes; ++c) {
            outputs[c] /= (double)n;                                    /* :215 */
            if (max_output < outputs[c]) { max_output = outputs[c]; best = c; }
        }
        out_label[i] = best;
    }
}

void fir_oracle_pnn_seq(const double* train, const int32_t* train_label, int64_t n, int d, int n_classes,
                        const double* avg, const double* q, int64_t nq, int32_t* out_label) {
    const int delta = 32;                                               /* delta_features_count, :182 */
    double var = 0.00002;                                               /* :230 */
    if (d > 2000) var /= 10;                                            /* :232-233 */
    double* dist = (double*)malloc(sizeof(double) * (size_t)n);
    double* outputs = (double*)malloc(sizeof(double) * (size_t)n_classes);
    char* check = (char*)malloc((size_t)n_classes);
    for (int64_t i = 0; i < nq; ++i) {
        int bestClass = -1;
        for (int64_t t = 0; t < n; ++t) dist[t] = 0;                    /* :237-240 */
        for (int c = 0; c < n_classes; ++c) { outputs[c] = 0; check[c] = 1; }
        const double den = (double)n;                                   /* :243 */
        for (int cur = 0; cur < d; cur += delta) {                      /* :245 */
            int max_fi = cur + delta;
            if (max_fi > d) max_fi = d;
            for (int c = 0; c < n_classes; ++c) if (check[c]) outputs[c] = 0;            /* :251 */
            for (int64_t t = 0; t < n; ++t) {                           /* class-major ⇒ the order of :249-264 */
                const int c = train_label[t];
                if (!check[c]) continue;
                for (int fi = cur; fi < max_fi; ++fi) {
                    double diff = train[t * d + fi] - avg[fi];          /* :258 */
                    double val = q[i * d + fi] - avg[fi];               /* :261 */
                    diff -= val;
                    dist[t] += diff * diff;                             /* :264 */
                }
                outputs[c] += exp(-dist[t] / (2 * var * (double)(size_t)max_fi));        /* :266 */
            }
            double max_output = -DBL_MAX;
            for (int c = 0; c < n_classes; ++c)
                if (check[c]) {
                    outputs[c] = outputs[c] / den;                      /* :268 */
                    if (max_output < outputs[c]) { max_output = outputs[c]; bestClass = c; }   /* :271-279 */
                }
            int variants = 0;
            const float thr = (float)(max_output / 1000000000);         /* :282, output_dividor = 1E9 (:185) */
            for (int c = 0; c < n_classes; ++c)
                if (check[c]) {
                    if (outputs[c] < thr) check[c] = 0; else ++variants;                  /* :283-289 */
                }
            if (variants == 1) break;                                   /* :291-292 */
        }
        out_label[i] = bestClass;
    }
    free(dist); free(outputs); free(check);
}

/* ------------------------------------------------------------------------------------------
 * FPNN — orthogonal-series PNN, classification.cpp:618-791; fasterlog2 :64-73
 * ------------------------------------------------------------------------------------------ */
float fir_oracle_fasterlog2(float x) {
    union { float f; uint32_t i; } vx = { x };
    union { uint32_t i; float f; } mx = { (vx.i & 0x007FFFFF) | (0x7e << 23) };
    float y = vx.i;
    y *= 1.0 / (1 << 23);
    return y - 124.22544637f - 1.498030302f * mx.f - 1.72587999f / (0.3520887068f + mx.f);
}

static double fpnn_normalize(double x, double avg, double sd, double scale) {     /* :633-654 */
    double val = (sd != 0) ? scale * (x - avg) / sd : 0;
    const double max_val = 0.5;
    if (val < -max_val) val = -max_val;
    else if (val > max_val) val = max_val;
    return val;
}

int fir_oracle_fpnn_J(int64_t n_train, int n_classes) {                            /* :666-673 */
    int J = (int)ceil(pow(1.0 * n_train / n_classes, 1.0 / 3));
    if (J <= 3) J = 3;
    return J;
}

/* train(): a[(fi*C + i)*(2J+1) + ...]; train rows are class-major (labels ascending) */
void fir_oracle_fpnn_train(const double* train, const int32_t* train_label, int64_t n, int d, int n_classes, const double* avg,
                           const double* sd, double scale, int J, double* a) {
    const double PI = atan(1.0) * 4;                                               /* :656 */
    int64_t* cls_begin = (int64_t*)calloc((size_t)n_classes + 1, sizeof(int64_t));
    for (int64_t t = 0; t < n; ++t) cls_begin[train_label[t] + 1]++;
    for (int c = 0; c < n_classes; ++c) cls_begin[c + 1] += cls_begin[c];
    const size_t J1 = (size_t)(2 * J + 1);
    for (size_t i = 0; i < (size_t)n_classes * d * J1; ++i) a[i] = 0;
    const size_t Jz = (size_t)J;
    for (int fi = 0; fi < d; ++fi)
        for (int i = 0; i < n_classes; ++i) {
            const size_t model_ind = ((size_t)fi * n_classes + i) * J1;
            a[model_ind] = 0.5;
            const int64_t cnt = cls_begin[i + 1] - cls_begin[i];
            const double cur_mult = 1.0 / cnt;
            for (int64_t t = cls_begin[i]; t < cls_begin[i + 1]; ++t) {
                const double val = fpnn_normalize(train[t * d + fi], avg[fi], sd[fi], scale);
                for (size_t j = 0; j < Jz; ++j) {
                    a[model_ind + 2 * j + 1] += cos(PI * (j + 1) * val) * cur_mult * (Jz - j) / (Jz * (Jz + 1));      /* :687 */
                    a[model_ind + 2 * j + 2] += sin(PI * (j + 1) * val) * cur_mult * (Jz - j) / (Jz * (Jz + 1));      /* :688 */
                }
            }
        }
    free(cls_begin);
}

/* predict_bf (:697-735) when sequential == 0, predict_sequentional (:736-791) otherwise */
void fir_oracle_fpnn_predict(const double* a, int J, int d, int n_classes, const double* avg, const double* sd, double scale,
                             const double* q, int64_t nq, int sequential, float output_ratio, int32_t* out_label) {
    const double PI = atan(1.0) * 4;
    const size_t J1 = (size_t)(2 * J + 1);
    const float output_delta = fir_oracle_fasterlog2(output_ratio);                /* ctor, :621 */
    float* outputs = (float*)malloc(sizeof(float) * (size_t)n_classes);
    int* check = (int*)malloc(sizeof(int) * (size_t)n_classes);
    double* cos_vals = (double*)malloc(sizeof(double) * (size_t)J);
    double* sin_vals = (double*)malloc(sizeof(double) * (size_t)J);
    for (int64_t qi = 0; qi < nq; ++qi) {
        const double* x = q + qi * d;
        int bestClass = -1;
        for (int i = 0; i < n_classes; ++i) { outputs[i] = 0; check[i] = 1; }
        const int delta = sequential ? 32 : d;                                     /* delta_features_count, :182 */
        for (int cur = 0; cur < d; cur += delta) {
            int max_fi = cur + delta;
            if (max_fi > d) max_fi = d;
            for (int fi = cur; fi < max_fi; ++fi) {
                const double val = fpnn_normalize(x[fi], avg[fi], sd[fi], scale);
                cos_vals[0] = cos(PI * val);
                sin_vals[0] = sin(PI * val);
                for (int j = 1; j < J; ++j) {
                    cos_vals[j] = cos_vals[j - 1] * cos_vals[0] - sin_vals[j - 1] * sin_vals[0];
                    sin_vals[j] = cos_vals[j - 1] * sin_vals[0] + sin_vals[j - 1] * cos_vals[0];
                }
                for (int i = 0; i < n_classes; ++i) {
                    if (!check[i]) continue;
                    const size_t model_ind = ((size_t)fi * n_classes + i) * J1;
                    double probab = a[model_ind];
                    for (int j = 0; j < J; ++j) probab += (a[model_ind + 2 * j + 1] * cos_vals[j] + a[model_ind + 2 * j + 2] * sin_vals[j]);
                    outputs[i] += fir_oracle_fasterlog2(probab);
                }
            }
            float max_output = -FLT_MAX;
            for (int i = 0; i < n_classes; ++i)
                if (check[i] && max_output < outputs[i]) { max_output = outputs[i]; bestClass = i; }
            if (!sequential) break;
            int variants = 0;
            const float thr = max_output + output_delta * max_fi;                  /* :777 */
            for (int i = 0; i < n_classes; ++i) {                                  /* :779-784: dropped classes are re-counted on stale outputs */
                if (outputs[i] < thr) check[i] = 0;
                else ++variants;
            }
            if (variants == 1) break;
        }
        out_label[qi] = bestClass;
    }
    free(outputs); free(check); free(cos_vals); free(sin_vals);
}

/* ------------------------------------------------------------------------------------------
 * PNN with clustering — per-class k-medoids, classification.cpp:321-388 (100 steps, medoid = the member with the smallest
 * summed distance to its cluster, distances on the RAW rows).  out_selected: positions in the class-major training order.
 * Returns the number selected, or -1 when a cluster runs empty (the reference then indexes with (size_t)-1).
 * ------------------------------------------------------------------------------------------ */
int64_t fir_oracle_kmedoids(const double* train, const int32_t* train_label, int64_t n, int d, int n_classes, int num_clusters,
                            int64_t* out_selected) {
    int64_t* cls_begin = (int64_t*)calloc((size_t)n_classes + 1, sizeof(int64_t));
    for (int64_t t = 0; t < n; ++t) cls_begin[train_label[t] + 1]++;
    for (int c = 0; c < n_classes; ++c) cls_begin[c + 1] += cls_begin[c];
    int64_t total = 0;
    for (int i = 0; i < n_classes; ++i) {
        const int64_t m = cls_begin[i + 1] - cls_begin[i];
        const double* rows = train + cls_begin[i] * d;
        if (m <= num_clusters) {                                                   /* :381-386 */
            for (int64_t j = 0; j < m; ++j) out_selected[total++] = cls_begin[i] + j;
            continue;
        }
        double* M = (double*)malloc(sizeof(double) * (size_t)m * m);
        for (int64_t u = 0; u < m; ++u)
            for (int64_t v = 0; v < m; ++v) {
                double dist = 0;
                for (int fi = 0; fi < d; ++fi) dist += (rows[u * d + fi] - rows[v * d + fi]) * (rows[u * d + fi] - rows[v * d + fi]);
                M[u * m + v] = dist / d;
            }
        int64_t* best = (int64_t*)malloc(sizeof(int64_t) * (size_t)m);
        int64_t* cent = (int64_t*)malloc(sizeof(int64_t) * (size_t)num_clusters);
        for (int c = 0; c < num_clusters; ++c) cent[c] = c;
        int dead = 0;
        for (int step = 0; step < 100 && !dead; ++step) {
            for (int64_t t = 0; t < m; ++t) {                                      /* :333-352 */
                best[t] = -1;
                double bestDist = DBL_MAX;
                for (int c = 0; c < num_clusters; ++c) {
                    const double dist = M[cent[c] * m + t];
                    if (dist < bestDist) { bestDist = dist; best[t] = c; }
                }
            }
            for (int c = 0; c < num_clusters; ++c) {                               /* :353-375 */
                double bestClustDist = DBL_MAX;
                cent[c] = -1;
                for (int64_t t = 0; t < m; ++t)
                    if (best[t] == c) {
                        double clustDist = 0;
                        for (int64_t t1 = 0; t1 < m; ++t1)
                            if (best[t1] == c) clustDist += M[t * m + t1];
                        if (clustDist < bestClustDist) { bestClustDist = clustDist; cent[c] = t; }
                    }
                if (cent[c] < 0) dead = 1;
            }
        }
        if (!dead)
            for (int c = 0; c < num_clusters; ++c) out_selected[total++] = cls_begin[i] + cent[c];
        free(M); free(best); free(cent);
        if (dead) { free(cls_begin); return -1; }
    }
    free(cls_begin);
    return total;
}

/* ------------------------------------------------------------------------------------------
 * Three-way-decision classifiers — ImageTesting.cpp:74-186 (conventional) and :188-288 (proposed,
 * CHECK_ALL_INSTANCES build).  last_feature is the reference's hard-coded 256 (:168, :221).
 * ------------------------------------------------------------------------------------------ */
static int desc_cmp(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x < y) - (x > y);
}

void fir_oracle_twd_conventional(int metric, const float* g, const int32_t* labels, int64_t n, int d, int n_classes,
                                 const float* q, int64_t nq, int type, double threshold, int feat_count, int last_feature,
                                 int32_t* out_idx, int32_t* out_class, uint8_t* out_unreliable) {
    double* distances = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double* probabs = (double*)malloc(sizeof(double) * (size_t)n_classes);
    (void)d;
    for (int64_t i = 0; i < nq; ++i) {
        const float* lhs = q + i * d;
        int bestInd = -1;
        double bestDist = 100000, secondBestDist = 100000;          /* :111 */
        double max_probab = 0, probab = 0;
        for (int c = 0; c < n_classes; ++c) probabs[c] = 0;
        for (int64_t j = 0; j < n; ++j) {                           /* :116-133 */
            distances[j] = fir_oracle_distance(metric, lhs, g + j * d, 0, feat_count);
            if (type == 0) {
                probab = exp(-distances[j] * 100);                  /* DIST_WEIGHT, :113,119 */
                if (probab > probabs[labels[j]]) probabs[labels[j]] = probab;
            }
            if (distances[j] < bestDist) {
                if (bestInd != -1 && labels[bestInd] != labels[j]) secondBestDist = bestDist;   /* previous best, not the true runner-up */
                bestDist = distances[j];
                bestInd = (int)j;
                if (type == 0) max_probab = probab;
            }
        }
        int reliable = 0;
        if (type == 0) {
            /* :143-150: nth_element(…, 5, greater) then the sum of the first five = the five largest posteriors.  The order
             * in which libstdc++ leaves them (and so the last bits of the sum) is unspecified; summed here largest first. */
            qsort(probabs, (size_t)n_classes, sizeof(double), desc_cmp);
            double sum = 0;
            for (int c = 0; c < 5; ++c) sum += probabs[c];
            max_probab /= sum;
            reliable = max_probab > threshold;
        } else if (type == 1) {
            reliable = (secondBestDist - bestDist) > threshold;     /* :160 */
        } else {
            reliable = (bestDist / secondBestDist) < threshold;     /* :163 */
        }
        if (!reliable) {                                            /* :166-181 */
            bestInd = -1;
            bestDist = 100000;
            for (int64_t j = 0; j < n; ++j) {
                /* double*int + float*int: the second product is single precision (:174-175) */
                const float rest = fir_oracle_distance(metric, lhs, g + j * d, feat_count, last_feature) * (float)(last_feature - feat_count);
                distances[j] = (distances[j] * feat_count + rest) / last_feature;
                if (distances[j] < bestDist) { bestDist = distances[j]; bestInd = (int)j; }
            }
        }
        if (out_unreliable) out_unreliable[i] = (uint8_t)!reliable;
        if (out_idx) out_idx[i] = bestInd;
        if (out_class) out_class[i] = bestInd >= 0 ? labels[bestInd] : -1;
    }
    free(distances); free(probabs);
}

void fir_oracle_twd_proposed(int metric, const float* g, const int32_t* labels, int64_t n, int d, const float* q, int64_t nq,
                             int feat_count, double th, int last_feature, int32_t* out_idx, int32_t* out_class, uint8_t* out_unreliable) {
    const double threshold = 1.0 / th;                              /* ctor, :191 */
    double* distances = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    char* check = (char*)malloc((size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < nq; ++i) {
        const float* lhs = q + i * d;
        int bestInd = -1;
        uint8_t unreliable = 0;
        for (int64_t j = 0; j < n; ++j) { distances[j] = 0; check[j] = 1; }
        for (int cur = 0; cur < last_feature; cur += feat_count) {  /* :223 */
            double bestDist = 100000;
            for (int64_t j = 0; j < n; ++j) {
                if (!check[j]) continue;
                distances[j] += fir_oracle_distance(metric, lhs, g + j * d, cur, cur + feat_count);   /* :243 */
                if (distances[j] < bestDist) { bestDist = distances[j]; bestInd = (int)j; }
            }
            int variants = 1;
            const double dist_threshold = bestDist * threshold;     /* :256 */
            const int bestClass = labels[bestInd];
            for (int64_t j = 0; j < n; ++j)
                if (check[j]) {
                    if (distances[j] > dist_threshold) check[j] = 0;
                    else if (labels[j] != bestClass) ++variants;
                }
            if (variants == 1) break;
            if (cur == 0) unreliable = 1;                           /* :281-282 */
        }
        if (out_unreliable) out_unreliable[i] = unreliable;
        if (out_idx) out_idx[i] = bestInd;
        if (out_class) out_class[i] = bestInd >= 0 ? labels[bestInd] : -1;
    }
    free(distances); free(check);
}

/* ------------------------------------------------------------------------------------------
 * DirectedEnumeration build — ann.cpp:270-348 (PIVOT branch), init :357-386, getThreshold :84-93
 * ------------------------------------------------------------------------------------------ */
static int fcmp(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

int fir_oracle_dem_build(int metric, const float* g, const int32_t* labels, int64_t n, int d, int pivot0,
                         float far_, float threshold_in, int keep_rows, int* np_out, int32_t* pivots,
                         float* P, float* min_other, float* threshold_out) {
    int np = (int)(n * 0.015);                                          /* :373 */
    if (np < 5) np = 5;                                                 /* :374-375 */
    if (np_out) *np_out = np;
    double* S = (double*)calloc((size_t)n, sizeof(double));             /* running form of :315-321 */
    float* row = (float*)malloc(sizeof(float) * (size_t)n);
    float* others = (float*)malloc(sizeof(float) * (size_t)np);
    pivots[0] = pivot0;
    for (int ii = 0; ii < np; ++ii) {
        int i = pivots[ii];
        int mostFarModel = -1;
        double maxFarDist = 0;
        float min_other_dist = FLT_MAX;
        for (int64_t j = 0; j < n; ++j) {
            float dd = fir_oracle_distance(metric, g + j * d, g + (int64_t)i * d, 0, d);   /* :309 lhs=db[j] rhs=db[i] */
            row[j] = dd;
            if (labels[i] != labels[j] && dd < min_other_dist) min_other_dist = dd;         /* :312-314 */
            if (j == i) S[j] = -1000000;                                                    /* :317-318 */
            else S[j] += dd;                                                                /* :320 */
            if (S[j] > maxFarDist) { maxFarDist = S[j]; mostFarModel = (int)j; }            /* :322-325 */
        }
        if (ii < keep_rows && P) memcpy(P + (size_t)ii * n, row, sizeof(float) * (size_t)n);
        others[ii] = min_other_dist;                                                        /* :327 */
        if (min_other) min_other[ii] = min_other_dist;
        if (ii < np - 1) pivots[ii + 1] = mostFarModel;                                     /* :328-330 */
    }
    float thr = threshold_in;
    if (!(threshold_in > 0)) {                                                              /* :340-342 */
        int ind = (int)(np * far_);                                                         /* :86 */
        qsort(others, (size_t)np, sizeof(float), fcmp);                                     /* nth_element value */
        thr = others[ind];
    }
    if (threshold_out) *threshold_out = thr;
    free(S); free(row); free(others);
    return np > 32 ? 32 : np;                                                               /* :332-333 */
}

/* ------------------------------------------------------------------------------------------
 * DirectedEnumeration::recognize — ann.cpp:416-507 (PIVOT: only the branch at :474-477 runs)
 * ------------------------------------------------------------------------------------------ */
static const float* g_sort_lik;
static int lik_cmp(const void* a, const void* b) {
    int x = *(const int*)a, y = *(const int*)b;
    if (g_sort_lik[x] < g_sort_lik[y]) return -1;
    if (g_sort_lik[y] < g_sort_lik[x]) return 1;
    return (x > y) - (x < y);             /* std::partial_sort is unstable; ties defined by index */
}

void fir_oracle_dem_search(int metric, const float* g, int64_t n, int d, const int32_t* pivots, int n_pivots,
                           const float* P, float threshold, int count_to_check, const float* q, int64_t nq,
                           int32_t* out_idx, float* out_dist, uint8_t* out_below, int32_t* out_evals) {
    int M = (count_to_check > 0 && count_to_check < n) ? count_to_check : (int)n;           /* ann.h:20-22 */
    float* lik = (float*)malloc(sizeof(float) * (size_t)n);
    int* li = (int*)malloc(sizeof(int) * (size_t)n);
    for (int64_t qi = 0; qi < nq; ++qi) {
        const float* qv = q + qi * d;
        int bestIndex = -1, below = 0, count = 0, start_index = 0;
        float bestDistance = FLT_MAX, tmpDist;
        for (int64_t i = 0; i < n; ++i) { lik[i] = 0; li[i] = (int)i; }                     /* :432-435 */
        for (int i = 0; i < n_pivots; ++i) {                                                /* :441 */
            int imageNum = pivots[i];
            tmpDist = fir_oracle_distance(metric, qv, g + (int64_t)imageNum * d, 0, d); ++count;   /* :391 */
            if (tmpDist < bestDistance) {
                bestDistance = tmpDist; bestIndex = imageNum;
                if (bestDistance < threshold) { below = 1; goto end; }                      /* :396-399 */
            }
            li[imageNum] = li[start_index];                                                 /* :444 */
            li[start_index++] = imageNum;                                                   /* :445 */
            const float* P_row = P + (size_t)n * i;
            for (int64_t ii = start_index; ii < n; ++ii) {                                  /* :453-461 */
                int nu = li[ii];
                float modelsDist = P_row[nu];
                if (modelsDist >= 0) { float tmp = tmpDist - modelsDist; lik[nu] += tmp * tmp; }
            }
        }
        if (M > start_index) {
            g_sort_lik = lik;
            qsort(li + start_index, (size_t)(n - start_index), sizeof(int), lik_cmp);       /* :469-470 */
        }
        while (count < M) {                                                                 /* :472-477 */
            int imageNum = li[start_index++];
            tmpDist = fir_oracle_distance(metric, qv, g + (int64_t)imageNum * d, 0, d); ++count;
            if (tmpDist < bestDistance) {
                bestDistance = tmpDist; bestIndex = imageNum;
                if (bestDistance < threshold) { below = 1; goto end; }
            }
        }
    end:
        out_idx[qi] = bestIndex;
        if (out_dist) out_dist[qi] = bestDistance;
        if (out_below) out_below[qi] = (uint8_t)below;
        if (out_evals) out_evals[qi] = count;
    }
    free(lik); free(li);
}
