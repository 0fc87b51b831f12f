/* TEST INFRASTRUCTURE ONLY — CPU restatement of the reference matching path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load libfir_oracle.so; the product (fast-image-recognition_b200/) never does.
 * Parity pin: checked bit-for-bit against oracle/_ref (the unmodified reference translation
 * units compiled from /root/reference) by tests/test_oracle_vs_ref.py, and against the frozen
 * vectors under tests/golden/ (generated from oracle/_ref by tests/golden/make_golden.py).
 * The reference itself ships no tests or golden vectors.
 */
#ifndef FIR_ORACLE_H
#define FIR_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { FIR_ORACLE_L2 = 0, FIR_ORACLE_CHI2 = 1, FIR_ORACLE_KL = 2 };

/* glibc 2.39 logf (sysdeps/ieee754/flt-32/e_logf.c; x86_64 ifunc variant __logf_fma when fma!=0,
 * the SSE2 build when fma==0) restated; the reference's KL branch calls it (db_features.cpp:34-36). */
float fir_oracle_logf(float x, int fma);
/* which variant the host libm resolves to: 1 fma, 0 sse2, -1 neither matched */
int fir_oracle_logf_host_variant(void);

/* feature_distance, db_features.cpp:22-42 */
float fir_oracle_distance(int metric, const float* lhs, const float* rhs, int start_pos, int end_pos);
/* loader normalisation, db_features.cpp:79-101 (zero |x|<1e-4, then L2 or L1 normalise) */
void fir_oracle_normalize_rows(int metric, float* rows, int64_t n, int d);

/* BruteForce::recognize ann.cpp:113-126 / recognize_image_bf db_features.cpp:319-335 */
double fir_oracle_bf(int metric, const float* g, int64_t n, int d, const float* q, int64_t nq,
                     int max_features, int nthreads, int32_t* out_idx, float* out_dist);
/* k lexicographically smallest (dist, j): k rounds of the reference argmin with removal */
double fir_oracle_topk(int metric, const float* g, int64_t n, int d, const float* q, int64_t nq, int k,
                       int nthreads, int32_t* out_idx, float* out_dist);
/* per-class nearest neighbour: for every class the reference argmin restricted to that class */
void fir_oracle_class_min(int metric, const float* g, const int32_t* labels, int64_t n, int d, int n_classes,
                          const float* q, int64_t nq, float* out_min, int32_t* out_arg);
/* PNN-style class scores over the fp32 divergence: score_c = (1/n) * sum_{j in c} exp(-dist_j / (2*var)), fp64 */
void fir_oracle_pnn_div(int metric, const float* g, const int32_t* labels, int64_t n, int d, int n_classes,
                        const float* q, int64_t nq, double var, double* out_scores, int32_t* out_label);

/* KNNClassifier::predict classification.cpp:116-170 and PNNClassifier::predict_bf :188-226 (fp64).
 * train rows are class-major in the order predict() walks training_set; avg = avgValues (:969-989). */
void fir_oracle_knn(const double* train, const int32_t* train_label, int64_t n, int d, int n_classes,
                    const double* avg, const double* q, int64_t nq, int K, int32_t* out_label);
void fir_oracle_pnn(const double* train, const int32_t* train_label, int64_t n, int d, int n_classes,
                    const double* avg, const double* q, int64_t nq, double* out_scores, int32_t* out_label);

/* PNNClassifier::predict_sequentional classification.cpp:228-295: 32-dimension chunks, classes whose score falls below
 * max/1e9 are dropped, stop when one class is left. */
void fir_oracle_pnn_seq(const double* train, const int32_t* train_label, int64_t n, int d, int n_classes,
                        const double* avg, const double* q, int64_t nq, int32_t* out_label);

/* FPNNClassifier (orthogonal-series PNN) classification.cpp:618-791, fasterlog2 :64-73.  train rows are the RAW rows in
 * class-major order; avg / sd = avgValues / stdValues of split_train_test (:969-989).  a has d*C*(2J+1) entries. */
float fir_oracle_fasterlog2(float x);
int fir_oracle_fpnn_J(int64_t n_train, int n_classes);
void fir_oracle_fpnn_train(const double* train, const int32_t* train_label, int64_t n, int d, int n_classes, const double* avg,
                           const double* sd, double scale, int J, double* a);
void fir_oracle_fpnn_predict(const double* a, int J, int d, int n_classes, const double* avg, const double* sd, double scale,
                             const double* q, int64_t nq, int sequential, float output_ratio, int32_t* out_label);
/* PNNwithClusteringClassifier::train classification.cpp:321-388: per-class k-medoids on the raw rows; positions in the
 * class-major training order; -1 when a cluster runs empty (undefined behaviour in the reference). */
int64_t fir_oracle_kmedoids(const double* train, const int32_t* train_label, int64_t n, int d, int n_classes, int num_clusters,
                            int64_t* out_selected);

/* ConventionalTWDClassifier::recognize ImageTesting.cpp:108-186 (type 0 Posteriors, 1 DistDiff, 2 DistRatio) and
 * ProposedTWDClassifier::recognize :207-288; last_feature = 256 in the reference.  out_unreliable = the query's
 * contribution to num_of_unreliable. */
void fir_oracle_twd_conventional(int metric, const float* g, const int32_t* labels, int64_t n, int d, int n_classes,
                                 const float* q, int64_t nq, int type, double threshold, int feat_count, int last_feature,
                                 int32_t* out_idx, int32_t* out_class, uint8_t* out_unreliable);
void fir_oracle_twd_proposed(int metric, const float* g, const int32_t* labels, int64_t n, int d, const float* q, int64_t nq,
                             int feat_count, double th, int last_feature, int32_t* out_idx, int32_t* out_class, uint8_t* out_unreliable);

/* DirectedEnumeration ctor + init, ann.cpp:270-348,357-386 (PIVOT build); getThreshold :84-93.
 * pivot0 replaces the first element of the reference's random_shuffle (:369).  keep_rows rows of the
 * pivot-distance matrix are written to P (keep_rows x n); np_out = max(5,(int)(n*0.015)) rows are walked. */
int fir_oracle_dem_build(int metric, const float* g, const int32_t* labels, int64_t n, int d, int pivot0,
                         float far_, float threshold_in, int keep_rows, int* np_out, int32_t* pivots,
                         float* P, float* min_other, float* threshold_out);
/* DirectedEnumeration::recognize, ann.cpp:416-507; candidate order = ascending (likelihood, index) */
void fir_oracle_dem_search(int metric, const float* g, int64_t n, int d, const int32_t* pivots, int n_pivots,
                           const float* P, float threshold, int count_to_check, const float* q, int64_t nq,
                           int32_t* out_idx, float* out_dist, uint8_t* out_below, int32_t* out_evals);

#ifdef __cplusplus
}
#endif
#endif
