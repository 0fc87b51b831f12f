/* fir_b200.h — C-ABI of the B200-native matching engine (libfir_b200.so).
 *
 * This is the drop-in boundary for the matching hot path of av-savchenko/fast-image-recognition's
 * qt_cpp recognizer.  The reference has no FFI layer; its path sits behind ordinary C++ declarations
 * (qt_cpp/db_features.h, qt_cpp/ann.h, and the Classifier shape inside qt_cpp/classification.cpp).
 * Each entry point below names the reference interface it replaces (paths relative to
 * /root/reference/).  Same-named C++ adapters over this C-ABI live in include/fir_b200_compat.hpp;
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns a fir_status (0 = ok) and records a
 *     message retrievable with fir_last_error_string() (thread-local).
 *   - `memspace` says whether the data pointers of THAT call are host (FIR_HOST: the call copies
 *     host<->device on the handle's stream and synchronises before returning) or device
 *     (FIR_DEVICE: fully asynchronous on the handle's stream, no synchronisation).
 *   - gallery/query rows are row-major fp32, D contiguous floats per row, "as loaded by
 *     db_features" (qt_cpp/db_features.cpp:79-101), i.e. already normalised by the caller or by
 *     fir_normalize_rows().
 *   - indices written to the caller are GLOBAL gallery indices (local row + index_offset) as int32;
 *     -1 means "no match" exactly as in the reference (qt_cpp/ann.cpp:100, db_features.cpp:322).
 *   - there is NO CPU fallback anywhere behind this interface: without a CUDA device every compute
 *     entry point fails with FIR_ERR_CUDA.
 */
#ifndef FIR_B200_H
#define FIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FIR_B200_VERSION 100

typedef enum fir_status {
    FIR_OK = 0,
    FIR_ERR_BAD_ARG = 1,
    FIR_ERR_CUDA = 2,
    FIR_ERR_OOM = 3,
    FIR_ERR_UNSUPPORTED = 4,
    FIR_ERR_INTERNAL = 5,
    FIR_ERR_NCCL = 6
} fir_status;

/* compile-time switch of the reference: USE_L2_DISTANCE (qt_cpp/db_features.h:12) and the
 * chi-square / KL `#if` (qt_cpp/db_features.cpp:30) become a run-time metric. */
typedef enum fir_metric { FIR_L2 = 0, FIR_CHI2 = 1, FIR_KL = 2 } fir_metric;
typedef enum fir_memspace { FIR_HOST = 0, FIR_DEVICE = 1 } fir_memspace;
/* which kernels serve fir_search_topk */
typedef enum fir_path {
    FIR_PATH_AUTO = 0,   /* L2: tensor-core candidates + exact rerank; chi2/KL: approximate tiles + exact rerank;
                            <= 8 queries: one-pass streaming kernels                                  */
    FIR_PATH_EXACT = 1,  /* exact CUDA-core tile kernel for every metric                              */
    FIR_PATH_TENSOR = 2, /* force the tcgen05 path (L2 only)                                          */
    FIR_PATH_APPROX = 3  /* chi2/KL: fast-intrinsic approximate tiles + exact rerank + certificate     */
} fir_path;

typedef struct fir_gallery fir_gallery;       /* vector<ImageInfo> dbImages + its ImagesDatabase      */
typedef struct fir_dem fir_dem;               /* DirectedEnumeration                                  */
typedef struct fir_classifier fir_classifier; /* the fp64 training-set state of classification.cpp    */

/* per-call counters of the last fir_search_topk on a gallery */
typedef struct fir_search_stats {
    int32_t path_used;        /* fir_path actually taken                                              */
    int32_t n_fallback;       /* queries the tensor path could not certify and re-ran exactly        */
    int32_t n_candidates;     /* exact reranked candidates per query (tensor path)                    */
    int32_t gpu_launches;     /* kernels launched by the call                                         */
    float approx_err_bound;   /* max certified |approx - true| squared-distance bound over queries    */
    float reserved;
} fir_search_stats;

const char* fir_last_error_string(void);
int fir_version(void);
int fir_device_count(int* count);
int fir_set_device(int device);

/* ---- data model -------------------------------------------------------------------------------
 * replaces: ImagesDatabase / ImageInfo (qt_cpp/db_features.h:14-29) as consumed through
 * `std::vector<ImageInfo>& dbImages` by ClassificationMethod (qt_cpp/ann.h:11,28).
 * The handle COPIES rows/labels to the device (the reference borrows them); the caller may free its
 * buffers afterwards.  index_offset is the global index of local row 0 (gallery row-sharding across
 * GPUs: one handle per process/GPU). */
int fir_gallery_create(const float* rows, const int32_t* labels, int64_t n, int32_t d, int32_t metric,
                       int32_t memspace, int64_t index_offset, fir_gallery** out);
int fir_gallery_destroy(fir_gallery* g);
int fir_gallery_set_stream(fir_gallery* g, void* cuda_stream);
int fir_gallery_info(const fir_gallery* g, int64_t* n, int32_t* d, int32_t* metric, int32_t* n_classes);
int fir_gallery_index_offset(const fir_gallery* g, int64_t* index_offset); /* global index of local row 0 (a loaded index carries its own) */
/* the class count defaults to max(label)+1 of THIS handle's rows; a row shard sets the global count so that per-class
 * outputs (fir_class_min, fir_pnn_scores) have the same width on every shard */
int fir_gallery_set_num_classes(fir_gallery* g, int32_t n_classes);

/* replaces: the normalisation loop of loadImages (qt_cpp/db_features.cpp:79-101): zero |x|<1e-4,
 * then x /= sqrtf(sum x*x) (L2) or x /= sum x (chi2/KL), fp32, sequential.  In place.
 * loadVideos (qt_cpp/video.cpp:64-84) is the same loop, except that its non-L2 build divides by sum x*x (no square root):
 * pass FIR_NORM_VIDEO_SUMSQ as `metric` for that variant. */
enum { FIR_NORM_VIDEO_SUMSQ = 3 };
int fir_normalize_rows(float* rows, int64_t n, int32_t d, int32_t metric, int32_t memspace, void* cuda_stream);

/* ---- brute force ------------------------------------------------------------------------------
 * replaces: BruteForce::recognize (qt_cpp/ann.cpp:113-126) and recognize_image_bf
 * (qt_cpp/db_features.cpp:319-335) looped over the query set by
 * ClassificationMethod::testSetRecognition (qt_cpp/ann.cpp:94-109).
 * Writes, per query, the k lexicographically smallest (feature_distance, gallery index) pairs —
 * k = 1 is the reference argmin (strict '<' ⇒ lowest index on ties, -1 if nothing is < 100000).
 * max_features > 0 restricts the distance to the first max_features dimensions and divides by it
 * (recognize_image_bf's prefix mode); 0 = all D — Euclidean prefixes run on the tensor path as well (the shadow's
 * k-blocks up to the prefix, prefix row norms).  out_idx/out_dist: nq x k. */
int fir_search_topk(fir_gallery* g, const float* queries, int64_t nq, int32_t k, int32_t max_features,
                    int32_t path, int32_t memspace, int32_t* out_idx, float* out_dist);
int fir_search_last_stats(const fir_gallery* g, fir_search_stats* stats);

/* measurement aid: when enabled, CUDA event pairs are recorded on the handle's stream around every launch
 * of the dominant kernels; fir_profile_read synchronises and returns their summed duration and count since
 * fir_profile_enable(g, 1) was last called. */
typedef enum fir_kernel {
    FIR_KERNEL_L2_CANDIDATES = 0,       /* tcgen05 candidate kernel, first pass (all queries)            */
    FIR_KERNEL_EXACT_TILES = 1,
    FIR_KERNEL_DEM_LIKELIHOOD = 2,
    FIR_KERNEL_L2_CANDIDATES_PASS2 = 3, /* same kernel, second pass over the few uncertified queries      */
    FIR_KERNEL_STREAM_DISTANCES = 4,    /* small-batch one-pass streaming kernel (HBM-bound)               */
    FIR_KERNEL_APPROX_TILES = 5,        /* chi2/KL approximate tile kernel                                  */
    /* phases of the tensor path around the candidate kernel (first pass unless noted) */
    FIR_PHASE_PACK_QUERIES = 6,         /* fp16 shadow + norms + residuals of the query block               */
    FIR_PHASE_PRUNE = 7,
    FIR_PHASE_RERANK = 8,               /* exact fp32 distances of the surviving candidates                 */
    FIR_PHASE_SELECT = 9,               /* top-k + certificate                                              */
    FIR_PHASE_PASS2 = 10,               /* the whole second pass (gather, pack, kernel, prune, rerank, select, scatter) */
    FIR_PHASE_EXACT_RERUN = 11,         /* exact CUDA-core windows for what no pass certified (normally empty) */
    FIR_PHASE_SEED = 12                 /* sample pass that seeds the first pass's list thresholds          */
} fir_kernel;
int fir_profile_enable(fir_gallery* g, int32_t on);
int fir_profile_read(fir_gallery* g, int32_t kernel, double* total_ms, int32_t* launches);

/* diagnostic: the candidate lists of the last tensor-path fir_search_topk on this gallery (valid until the
 * next call): nq x n_slots x R local indices (-1 = empty), their tensor-core approximate squared distances
 * and their exact fp32 feature_distance values.  Pass NULL arrays to query n_slots/R first. */
int fir_debug_tensor_candidates(fir_gallery* g, int32_t* n_slots, int32_t* R, int32_t* idx, float* approx, float* exact);
/* Host-side self check of the tensor path's work partition (no device needed): 0 when every (query block, gallery tile) item of
 * an nq x n search over n_sm SMs (ctas per work unit = 1 or 2; row_bytes = bytes of one shadow row, decides the phased form of the
 * remainder for galleries larger than L2) is covered exactly once and the candidate slots of a query block are distinct. */
int fir_debug_partition_check(int64_t nq, int64_t n, int n_sm, int ctas, int64_t row_bytes, int* n_phases, int* n_slots);

/* replaces: ImageInfo::distance / feature_distance (qt_cpp/db_features.h:24-26, db_features.cpp:22-42)
 * for explicit (query, gallery index) pairs: cand_idx is nq x r LOCAL row indices (-1 = skip),
 * out_dist nq x r.  gallery_is_lhs != 0 evaluates feature_distance(gallery, query) — the operand
 * order of the DEM build (qt_cpp/ann.cpp:309); it only matters for KL. */
int fir_pair_distances(fir_gallery* g, const float* queries, int64_t nq, const int32_t* cand_idx, int32_t r,
                       int32_t gallery_is_lhs, int32_t memspace, float* out_dist);

/* per-class nearest neighbour: out_min/out_arg are nq x n_classes; classes without a match get
 * (100000, -1).  (Fused class reduction of the same distance tiles; no reference counterpart beyond
 * the argmin above.) */
int fir_class_min(fir_gallery* g, const float* queries, int64_t nq, int32_t memspace, float* out_min,
                  int32_t* out_arg);

/* PNN-style class scores over the gallery's fp32 divergence (the Parzen reducer of
 * PNNClassifier::predict_bf, qt_cpp/classification.cpp:213-225, with the divergence swapped in):
 * score[c] = (1/n_total) * sum_{j in class c} exp(-dist(q, x_j) / (2 * var)), fp64; label = argmax with
 * strict '<' (lowest class on ties).  n_total <= 0 means the handle's own row count (use the global
 * count when the gallery is sharded; the caller then sums scores across shards). */
int fir_pnn_scores(fir_gallery* g, const float* queries, int64_t nq, double var, int64_t n_total,
                   int32_t memspace, double* out_scores, int32_t* out_label);

/* k-way merge of per-shard top-k lists (multi-GPU candidate merge after an all-gather):
 * parts_* are n_parts x nq x k, each list sorted by (dist, idx), idx = -1 marks an empty slot;
 * out_* are nq x k.  Device pointers only. */
int fir_merge_topk(const float* parts_dist, const int32_t* parts_idx, int32_t n_parts, int64_t nq, int32_t k,
                   float* out_dist, int32_t* out_idx, void* cuda_stream);

/* ---- sequential three-way decisions (TWD) ---------------------------------------------------
 * replaces: ConventionalTWDClassifier::recognize (qt_cpp/ImageTesting.cpp:108-186) and
 * ProposedTWDClassifier::recognize (:207-288, the CHECK_ALL_INSTANCES build) over a gallery handle (train(&dbImages), :40).
 *   type            FIR_TWD_POSTERIORS / FIR_TWD_DIST_DIFF / FIR_TWD_DIST_RATIO  (TWD_Type, :76)
 *   threshold       the constructor's `th` (the proposed classifier applies 1/th itself, :191)
 *   feat_count      reduced_features_count (:77 default 64; 32 or 64 for the proposed classifier, :533-534)
 *   last_feature    256 in the reference (:168, :221)
 *   out_index       bestInd (global gallery index, -1 = none); out_label = its class (what recognize() returns);
 *   out_unreliable  the query's contribution to num_of_unreliable (:33, :167, :282).  Any output may be NULL. */
enum { FIR_TWD_POSTERIORS = 0, FIR_TWD_DIST_DIFF = 1, FIR_TWD_DIST_RATIO = 2 };
int fir_twd_conventional(fir_gallery* g, const float* queries, int64_t nq, int32_t type, double threshold, int32_t feat_count,
                         int32_t last_feature, int32_t memspace, int32_t* out_index, int32_t* out_label, uint8_t* out_unreliable);
int fir_twd_proposed(fir_gallery* g, const float* queries, int64_t nq, int32_t feat_count, double threshold, int32_t last_feature,
                     int32_t memspace, int32_t* out_index, int32_t* out_label, uint8_t* out_unreliable);

/* ---- fp64 kNN / PNN ---------------------------------------------------------------------------
 * replaces: the file-scope training state of qt_cpp/classification.cpp:53-62 as left by
 * split_train_test (:942-990) — training rows in class-major order (the order predict() walks
 * training_set), their class labels, and avgValues — plus KNNClassifier::predict (:116-170) and
 * PNNClassifier::predict_bf (:188-226). */
int fir_classifier_create(const double* train_rows, const int32_t* train_labels, int64_t n, int32_t d,
                          int32_t n_classes, const double* avg, fir_classifier** out);
int fir_classifier_destroy(fir_classifier* c);
int fir_classifier_knn(fir_classifier* c, const double* queries, int64_t nq, int32_t K, int32_t* out_label);
int fir_classifier_pnn(fir_classifier* c, const double* queries, int64_t nq, double* out_scores /* nq x C or NULL */,
                       int32_t* out_label);
/* The same two with an explicit memory space for the queries and outputs (FIR_DEVICE: asynchronous on the classifier's
 * stream).  The Parzen sums are fused into the epilogue of the fp64 distance tiles — the Q x N distance matrix is only
 * materialised for the kNN walk.  fir_classifier_profile(c, on, &ms, &launches) returns the summed CUDA-event duration of the
 * distance kernel since profiling was last switched on (NULL outputs: just switch). */
int fir_classifier_knn_ex(fir_classifier* c, const double* queries, int64_t nq, int32_t K, int32_t memspace, int32_t* out_label);
int fir_classifier_pnn_ex(fir_classifier* c, const double* queries, int64_t nq, int32_t memspace, double* out_scores, int32_t* out_label);
int fir_classifier_set_stream(fir_classifier* c, void* cuda_stream);
int fir_classifier_profile(fir_classifier* c, int32_t on, double* total_ms, int32_t* launches);
/* replaces: PNNClassifier::predict_sequentional (qt_cpp/classification.cpp:228-295; PNNClassifier(bf=false)): 32-dimension
 * chunks, classes scoring below max/1e9 are dropped, stop when one class is left.   */
int fir_classifier_pnn_sequential(fir_classifier* c, const double* queries, int64_t nq, int32_t* out_label);

/* ---- directed enumeration ---------------------------------------------------------------------
 * replaces: DirectedEnumeration (qt_cpp/ann.h:61-100): ctor + init (qt_cpp/ann.cpp:270-348,357-386),
 * getThreshold (:84-93), recognize (:416-507), setImageCountToCheck (qt_cpp/ann.h:20-22). */
typedef struct fir_dem_params {
    int32_t pivot0;          /* first pivot (the reference takes the head of an unseeded random_shuffle,
                                ann.cpp:366-369); < 0 ⇒ derived from `seed`                              */
    uint32_t seed;
    float false_accept_rate; /* 0.01f, ann.h:64                                                          */
    float threshold;         /* > 0 overrides the FAR quantile, ann.cpp:277-279,340                      */
    int32_t max_chain;       /* pivot-chain rows to walk; <= 0 ⇒ the reference's max(5,(int)(n*0.015))  */
    int32_t max_pivots;      /* pivots kept for search; <= 0 ⇒ 32, ann.cpp:332-333                       */
} fir_dem_params;

int fir_dem_build(fir_gallery* g, const fir_dem_params* params, fir_dem** out);
/* adopt an existing build (pivot list, n_pivots x n pivot-distance rows on the host, threshold) instead of running the
 * chain: the at-scale recipe of SURVEY.md §8(c), persisted indices, and parity tests of the search alone. */
int fir_dem_from_state(fir_gallery* g, const int32_t* pivots, int32_t n_pivots, const float* P, float threshold, fir_dem** out);
int fir_dem_destroy(fir_dem* dem);
int fir_dem_info(const fir_dem* dem, int32_t* n_pivots, int32_t* chain_rows, float* threshold);
/* counters of the last fir_dem_search: kernels launched; whether the first round runs as a tensor-core brute force in
 * pivot space (galleries of >= 8192 rows with 32 pivots); and, while fir_profile_enable is on for the gallery, the summed
 * duration / count of that round's tcgen05 candidate kernel since profiling was enabled. */
int fir_dem_search_stats(fir_dem* dem, int32_t* gpu_launches, int32_t* tensor_round, double* candidates_kernel_ms, int32_t* candidates_kernel_launches);
int fir_dem_get_pivots(const fir_dem* dem, int32_t* out_pivots /* n_pivots */);
int fir_dem_get_pivot_matrix(const fir_dem* dem, float* out_P /* n_pivots x n, host */);
int fir_dem_get_min_other(const fir_dem* dem, float* out /* chain_rows, host */);
/* out_idx: local row (+index_offset) or -1; out_dist = bestDistance; out_below = isFoundLessThreshold;
 * out_evals = distanceCalcCount.  count_to_check follows setImageCountToCheck. */
int fir_dem_search(fir_dem* dem, const float* queries, int64_t nq, int32_t count_to_check, int32_t memspace,
                   int32_t* out_idx, float* out_dist, uint8_t* out_below, int32_t* out_evals);

/* replaces: PNNwithClusteringClassifier (qt_cpp/classification.cpp:311-428).  train() (:321-388, per-class k-medoids over the
 * RAW training rows, class-major) is fir_kmedoids_select: out_selected receives the kept rows as positions in the given
 * order (at most n), *out_count how many.  predict() (:389-428) is the Parzen PNN over the kept rows divided by the FULL
 * training-set size: fir_classifier_create over the selected rows, fir_classifier_set_total(c, n), fir_classifier_pnn. */
int fir_kmedoids_select(const double* train_rows, const int32_t* train_labels, int64_t n, int32_t d, int32_t n_classes,
                        int32_t num_clusters, int64_t* out_selected, int64_t* out_count);
int fir_classifier_set_total(fir_classifier* c, int64_t n_total);

/* ---- FPNN: orthogonal-series PNN -------------------------------------------------------------
 * replaces: FPNNClassifier (qt_cpp/classification.cpp:618-791): train() (:658-695) at creation, predict_bf (:697-735) and
 * predict_sequentional (:736-791).  train_rows: RAW rows in class-major order; avg / std: avgValues / stdValues of
 * split_train_test (:969-989); features_scale and output_ratio are the constructor's arguments (:620). */
typedef struct fir_fpnn fir_fpnn;
int fir_fpnn_create(const double* train_rows, const int32_t* train_labels, int64_t n, int32_t d, int32_t n_classes,
                    const double* avg, const double* std, double features_scale, fir_fpnn** out);
int fir_fpnn_destroy(fir_fpnn* f);
int fir_fpnn_info(const fir_fpnn* f, int32_t* J, int64_t* n_coefficients);
int fir_fpnn_get_coefficients(const fir_fpnn* f, double* out_a /* d x C x (2J+1), host */);
int fir_fpnn_predict(fir_fpnn* f, const double* queries, int64_t nq, int32_t sequential, float output_ratio, int32_t* out_label);

/* ---- persisted index --------------------------------------------------
 * The reference keeps nothing on disk between runs: the gallery is re-parsed from the text features file
 * (qt_cpp/db_features.cpp:44-116) and DirectedEnumeration is rebuilt by its constructor (qt_cpp/ann.cpp:270-348).
 * One binary file holds the packed gallery (rows, labels, metric, class count, index offset) and, when `dem` is given,
 * its pivots, pivot-distance rows and threshold; a checksum guards the payload.  fir_index_load returns new handles
 * (the DEM handle borrows the gallery handle: destroy it first); *out_dem is NULL when the file carries no DEM state
 * or out_dem itself is NULL. */
int fir_index_save(const fir_gallery* g, const fir_dem* dem /* or NULL */, const char* path);
int fir_index_load(const char* path, fir_gallery** out_gallery, fir_dem** out_dem /* or NULL */);

/* ---- multi-GPU: gallery row shards + NCCL ----------------------------------------
 * The reference is single-device; its callers (testANN qt_cpp/ann.cpp:52-69, testSetRecognition :94-109) walk ONE
 * std::vector<ImageInfo>.  Here that vector is cut into contiguous row shards, one per GPU (global indices through
 * index_offset), queries are replicated, every GPU runs the complete single-GPU pipeline on its shard and ONE exchange
 * follows: top-k lists packed into 64-bit (ordered dist, idx) keys → ncclAllGather → k-way merge;
 * per-class minima → ncclAllReduce(min) on the same keys; PNN partial sums → ncclAllReduce(sum, fp64).
 * The answers equal the single-gallery ones bit for bit (ties are broken by GLOBAL index); fp64 PNN sums within 1e-5
 * relative (the summation order across shards differs).  NCCL is dlopen'ed on first use.
 *
 * (1) one process per GPU: every rank creates its own shard handle (fir_gallery_create with index_offset = first row)
 *     and joins a communicator; rank 0 obtains the id and the launcher distributes the 128 bytes. */
typedef struct fir_comm fir_comm;
#define FIR_COMM_ID_BYTES 128
int fir_comm_unique_id(void* id_out /* FIR_COMM_ID_BYTES */);
int fir_comm_init_rank(const void* id, int32_t rank, int32_t world, fir_comm** out); /* on the current device; world = 1 needs no id */
int fir_comm_destroy(fir_comm* c);
int fir_comm_info(const fir_comm* c, int32_t* rank, int32_t* world, int32_t* nccl_version);
/* Collective: every rank calls with the same queries (FIR_DEVICE: replicated device batch, asynchronous on the shard's
 * stream; FIR_HOST: the same host batch on every rank — each rank uploads 1/world of the rows over its own PCIe link,
 * an NVLink all-gather assembles the batch, the call synchronises) and every rank receives the merged result. */
int fir_shard_search_topk(fir_gallery* shard, fir_comm* c, const float* queries, int64_t nq, int32_t k, int32_t path, int32_t memspace,
                          int32_t* out_idx, float* out_dist);
int fir_shard_class_min(fir_gallery* shard, fir_comm* c, const float* queries, int64_t nq, int32_t memspace, float* out_min, int32_t* out_arg);
int fir_shard_pnn_scores(fir_gallery* shard, fir_comm* c, const float* queries, int64_t nq, double var, int64_t n_total, int32_t memspace,
                         double* out_scores, int32_t* out_label);
/* Directed enumeration over row shards (collective): ONE DirectedEnumeration over the whole gallery — the same pivots,
 * threshold and candidate order as a single-gallery build (qt_cpp/ann.cpp:270-507).  Rank r holds the pivot-distance columns
 * of its own rows; per chain step the pivot's row travels by all-reduce and the (far-sum maximum, row, minimum other-class
 * distance) triples by all-gather; in the search every rank walks its rows and each round's candidates (likelihood, global
 * row, exact distance) are all-gathered and folded in likelihood order identically on every rank.  fir_dem_search /
 * fir_dem_info / fir_dem_get_* on the returned handle are collective calls too; fir_dem_get_pivot_matrix returns the local
 * columns (n_pivots x shard rows); indices are global. */
int fir_shard_dem_build(fir_gallery* shard, fir_comm* c, int64_t n_total, const fir_dem_params* params, fir_dem** out);
/* (2) ONE process driving n_gpus devices (<= 0: all visible): rows/labels are the whole class-major gallery in host
 *     memory; shard r = rows [n*r/G, n*(r+1)/G) on device r; ncclCommInitAll.  Host buffers in and out. */
typedef struct fir_sharded fir_sharded;
int fir_sharded_create(const float* rows, const int32_t* labels, int64_t n, int32_t d, int32_t metric, int32_t n_gpus, fir_sharded** out);
int fir_sharded_destroy(fir_sharded* s);
int fir_sharded_info(const fir_sharded* s, int32_t* n_gpus, int64_t* n, int32_t* d, int32_t* n_classes);
int fir_sharded_shard(fir_sharded* s, int32_t rank, fir_gallery** shard, fir_comm** comm); /* borrowed handles of shard `rank` */
int fir_sharded_search_topk(fir_sharded* s, const float* queries, int64_t nq, int32_t k, int32_t path, int32_t* out_idx, float* out_dist);
int fir_sharded_class_min(fir_sharded* s, const float* queries, int64_t nq, float* out_min, int32_t* out_arg);
int fir_sharded_pnn_scores(fir_sharded* s, const float* queries, int64_t nq, double var, double* out_scores, int32_t* out_label);
/* DirectedEnumeration over the shards of a fir_sharded gallery (one host thread per GPU runs the collective rank-level calls):
 * same pivots, threshold and answers as ONE index over the whole gallery; global indices. */
typedef struct fir_sharded_dem fir_sharded_dem;
int fir_sharded_dem_build(fir_sharded* s, const fir_dem_params* params, fir_sharded_dem** out);
int fir_sharded_dem_destroy(fir_sharded_dem* d);
int fir_sharded_dem_info(const fir_sharded_dem* d, int32_t* n_pivots, int32_t* chain_rows, float* threshold);
int fir_sharded_dem_get_pivots(const fir_sharded_dem* d, int32_t* out_pivots /* n_pivots */);
int fir_sharded_dem_search(fir_sharded_dem* d, const float* queries, int64_t nq, int32_t count_to_check, int32_t* out_idx, float* out_dist,
                           uint8_t* out_below, int32_t* out_evals);

/* ---- synthetic workloads (bench / tests; no reference counterpart) ----------------------------------
 * Rows [row_lo, row_lo + n_rows) of the gallery (role 0) or query (role 1) matrix of a BASELINE.json config, generated on
 * the device by a counter-based Philox4x32-10 keyed by (seed, row, column) — csrc/synth_common.h — so that every rank of a
 * sharded run materialises its own rows of the SAME gallery and the host can regenerate identical bits
 * (libfir_synth_host.so: fir_synth_rows_host, same arguments + a thread count).  Values are class centroid + sigma * noise,
 * optionally ReLU'd; the caller applies fir_normalize_rows afterwards.  Gallery labels are class-major equal blocks
 * ((row * n_classes) / n_total), query labels pseudo-random.  Device pointers; out_labels may be NULL. */
int fir_synth_rows(float* out_rows, int32_t* out_labels, int64_t row_lo, int64_t n_rows, int64_t n_total, int32_t d,
                   int32_t n_classes, int32_t role, uint32_t seed, float sigma, int32_t relu, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* FIR_B200_H */
