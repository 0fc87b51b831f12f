// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or executed from the product
// path (fast-image-recognition_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library this file is built into.
//
// C-ABI driver around the UNMODIFIED reference translation units
//   /root/reference/qt_cpp/db_features.cpp, ann.cpp, classification.cpp
// compiled where they lie (see oracle/Makefile).  Everything numeric below is done by the
// reference's own functions; this file only (a) packs/unpacks flat arrays into the
// reference's data model (db_features.h:14-29), (b) fans queries out over std::threads with
// one matcher object per thread (the reference objects are not re-entrant, ann.h:29-31), and
// (c) reads/injects DirectedEnumeration's private state for the at-scale recipe in
// SURVEY.md §8(c).
#include <vector>
#include <string>
#include <unordered_map>
#include <map>
#include <thread>
#include <chrono>
#include <cstring>
#include <cstdio>
#include <sstream>
#include <fstream>
#include <set>
#include <locale>

// the driver needs DirectedEnumeration's / ClassificationMethod's internals (ann.h:26-31,92-99)
#define private public
#define protected public
#include "ann.h"
#undef private
#undef protected

// kNN / PNN and their file-scope state live in an anonymous namespace inside classification.cpp
// (:53-62), so that file is textually included here (only in the L2 variant; it has no metric).
#ifdef FIR_REF_WITH_CLASSIFICATION
#define main fir_ref_unused_main
#define private public      /* FPNNClassifier::a / J and PNNwithClusteringClassifier::clustered_training_set are read below */
#define protected public
#include "classification.cpp"
#undef private
#undef protected
#undef main
#endif

extern "C" {
int fir_ref_features_count = 1536;               // db.h:86 default
const char* fir_ref_features_file = "features.txt";
}

namespace {
struct Silence {                                  // the reference prints to cout from ctors
    std::ostringstream sink;          // declared first: must be constructed before its rdbuf is installed
    std::streambuf* old;
    Silence() : sink(), old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Silence() { std::cout.rdbuf(old); }
};

struct PackedDb {
    std::vector<FeaturesVector> rows;
    std::vector<ImageInfo> infos;
    PackedDb(const float* x, const int* labels, long n, int d) {
        rows.resize(n);
        infos.reserve(n);
        for (long j = 0; j < n; ++j) {
            rows[j].assign(x + j * (long)d, x + (j + 1) * (long)d);
            infos.push_back(ImageInfo(labels ? labels[j] : 0, (int)j, rows[j]));
        }
    }
};

template <typename F> double run_sharded(long nq, int nthreads, F body) {
    if (nthreads < 1) nthreads = 1;
    auto t1 = std::chrono::high_resolution_clock::now();
    if (nthreads == 1) {
        body(0, 0, nq);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) {
            long lo = nq * t / nthreads, hi = nq * (t + 1) / nthreads;
            th.emplace_back([=]() { body(t, lo, hi); });
        }
        for (auto& t : th) t.join();
    }
    auto t2 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double>(t2 - t1).count();
}
}  // namespace

extern "C" {

int fir_ref_metric(void) {
#ifdef USE_L2_DISTANCE
    return 0;
#elif defined(FIR_REF_KL)
    return 2;
#else
    return 1;
#endif
}

void fir_ref_set_dim(int d) { fir_ref_features_count = d; }

// feature_distance, db_features.cpp:22-42
float fir_ref_feature_distance(const float* l, const float* r, int d, int start, int end) {
    FeaturesVector a(l, l + d), b(r, r + d);
    return feature_distance(a, b, start, end);
}

// ---- data set: loadImages (db_features.cpp:44-116) + class filter (ann.cpp:34-37) +
//      getTrainingAndTestImages (db_features.cpp:117-162) ---------------------------------
struct fir_ref_dataset {
    ImagesDatabase all, kept;
    std::vector<ImageInfo> db, test;
    int d;
};

fir_ref_dataset* fir_ref_dataset_load(const char* path, int d, unsigned seed, int randomize) {
    Silence s;
    fir_ref_set_dim(d);
    fir_ref_dataset* h = new fir_ref_dataset();
    h->d = d;
    std::unordered_map<std::string, int> person2index;
    loadImages(h->all, path, person2index);
    for (auto& f : h->all)
        if (f.size() > 1) h->kept.push_back(f);
    srand(seed);
    getTrainingAndTestImages(h->kept, h->db, h->test, randomize != 0);
    return h;
}
long fir_ref_dataset_count(fir_ref_dataset* h, int which) { return which == 0 ? (long)h->db.size() : (long)h->test.size(); }
void fir_ref_dataset_get(fir_ref_dataset* h, int which, float* rows, int* labels, int* index_in_db) {
    std::vector<ImageInfo>& v = which == 0 ? h->db : h->test;
    for (size_t j = 0; j < v.size(); ++j) {
        std::memcpy(rows + j * (size_t)h->d, v[j].features.data(), sizeof(float) * h->d);
        labels[j] = v[j].classNo;
        if (index_in_db) index_in_db[j] = v[j].indexInDatabase;
    }
}
void fir_ref_dataset_free(fir_ref_dataset* h) { delete h; }

// ---- brute force: BruteForce::recognize (ann.cpp:113-126) or, with max_features>0,
//      recognize_image_bf (db_features.cpp:319-335) ------------------------------------------
double fir_ref_bf_search(const float* g, const int* glabels, long n, int d, const float* q, long nq,
                         int nthreads, int max_features, int* out_idx, float* out_dist) {
    Silence s;
    fir_ref_set_dim(d);
    PackedDb db(g, glabels, n, d);
    return run_sharded(nq, nthreads, [&](int, long lo, long hi) {
        BruteForce bf(db.infos);
        for (long i = lo; i < hi; ++i) {
            FeaturesVector qv(q + i * (long)d, q + (i + 1) * (long)d);
            ImageInfo qi(-1, -1, qv);
            int best = max_features > 0 ? recognize_image_bf(db.infos, qi, max_features) : bf.recognize(qi);
            out_idx[i] = best;
            if (out_dist)
                out_dist[i] = best < 0 ? 0.f
                              : (max_features > 0 ? qi.distance(db.infos[best], 0, max_features) : qi.distance(db.infos[best]));
        }
    });
}

// all Q x N distances through ImageInfo::distance (db_features.h:24-26); small cases only
void fir_ref_all_distances(const float* g, long n, int d, const float* q, long nq, int gallery_is_lhs, float* out) {
    fir_ref_set_dim(d);
    PackedDb db(g, 0, n, d);
    for (long i = 0; i < nq; ++i) {
        FeaturesVector qv(q + i * (long)d, q + (i + 1) * (long)d);
        ImageInfo qi(-1, -1, qv);
        for (long j = 0; j < n; ++j) out[i * n + j] = gallery_is_lhs ? db.infos[j].distance(qi) : qi.distance(db.infos[j]);
    }
}

// ---- DirectedEnumeration (ann.cpp:270-507) --------------------------------------------------
struct fir_ref_dem {
    PackedDb* db;
    std::vector<ImageInfo>* view;       // what DirectedEnumeration::dbImages refers to
    DirectedEnumeration* dem;
    int d;
    long np_built;                      // rows of P_matrix actually allocated by the ctor
};

fir_ref_dem* fir_ref_dem_create(const float* g, const int* glabels, long n, int d, unsigned seed,
                                float far_, float threshold, int count_to_check, double* build_seconds) {
    Silence s;
    fir_ref_set_dim(d);
    fir_ref_dem* h = new fir_ref_dem();
    h->d = d;
    h->db = new PackedDb(g, glabels, n, d);
    h->view = &h->db->infos;
    srand(seed);                        // libstdc++ random_shuffle draws from rand() (ann.cpp:369)
    auto t1 = std::chrono::high_resolution_clock::now();
    h->dem = new DirectedEnumeration(*h->view, far_, threshold, count_to_check);
    auto t2 = std::chrono::high_resolution_clock::now();
    if (build_seconds) *build_seconds = std::chrono::duration<double>(t2 - t1).count();
    int np = (int)(n * 0.015);
    if (np < 5) np = 5;
    h->np_built = np;
    return h;
}

// At-scale recipe: run the verbatim ctor on a tiny stand-in gallery, then grow the
// referenced vector to the full gallery and overwrite the private build state with a caller-supplied
// build (pivots, P rows, threshold).  recognize() that then runs is the verbatim reference code.
fir_ref_dem* fir_ref_dem_create_injected(const float* g, const int* glabels, long n, int d,
                                         const int* pivots, int n_pivots, const float* P /*n_pivots x n*/,
                                         float threshold) {
    Silence s;
    fir_ref_set_dim(d);
    if (n < 8) return 0;
    fir_ref_dem* h = new fir_ref_dem();
    h->d = d;
    h->db = new PackedDb(g, glabels, n, d);
    h->view = new std::vector<ImageInfo>(h->db->infos.begin(), h->db->infos.begin() + 8);
    srand(1);
    h->dem = new DirectedEnumeration(*h->view, 0.01f, 1.0f, 0);
    *h->view = std::vector<ImageInfo>();
    h->view->reserve(n);
    for (long j = 0; j < n; ++j) h->view->push_back(h->db->infos[j]);
    DirectedEnumeration* m = h->dem;
    delete[] m->P_matrix;
    delete[] m->likelihoods;
    delete[] m->likelihood_indices;
    m->P_matrix = new DirectedEnumeration::ImageDist[(size_t)n_pivots * n];
    for (int i = 0; i < n_pivots; ++i)
        for (long j = 0; j < n; ++j) m->P_matrix[(size_t)i * n + j] = DirectedEnumeration::ImageDist(P[(size_t)i * n + j], (int)j);
    m->likelihoods = new float[n];
    m->likelihood_indices = new int[n];
    m->startIndices.assign(pivots, pivots + n_pivots);
    m->threshold = threshold;
    m->imageCountToCheck = (int)n;
    h->np_built = n_pivots;
    return h;
}

int fir_ref_dem_num_pivots(fir_ref_dem* h) { return (int)h->dem->startIndices.size(); }
float fir_ref_dem_threshold(fir_ref_dem* h) { return h->dem->threshold; }
void fir_ref_dem_get_pivots(fir_ref_dem* h, int* out) {
    for (size_t i = 0; i < h->dem->startIndices.size(); ++i) out[i] = h->dem->startIndices[i];
}
// rows [0, n_rows) of the pivot-distance matrix, n_rows <= rows built
void fir_ref_dem_get_P(fir_ref_dem* h, int n_rows, float* out) {
    long n = (long)h->view->size();
    for (int i = 0; i < n_rows; ++i)
        for (long j = 0; j < n; ++j) out[(size_t)i * n + j] = h->dem->P_matrix[(size_t)i * n + j].dist;
}

// DirectedEnumeration::recognize over a query set (ann.cpp:416-507); single matcher ⇒ serial, or
// nthreads>1 ⇒ per-thread clones of the scratch arrays (the build state is shared read-only).
double fir_ref_dem_search(fir_ref_dem* h, const float* q, long nq, int count_to_check, int nthreads,
                          int* out_idx, float* out_dist, unsigned char* out_below, int* out_evals) {
    Silence s;
    fir_ref_set_dim(h->d);
    int d = h->d;
    long n = (long)h->view->size();
    h->dem->setImageCountToCheck(count_to_check);   // ann.h:20-22
    if (nthreads < 1) nthreads = 1;
    // per-thread shallow clones: own scratch + counters, shared P_matrix (read-only during search)
    std::vector<DirectedEnumeration*> clones(nthreads, (DirectedEnumeration*)0);
    std::vector<std::vector<char> > storage(nthreads);
    clones[0] = h->dem;
    std::vector<float*> lik(nthreads, (float*)0);
    std::vector<int*> lidx(nthreads, (int*)0);
    for (int t = 1; t < nthreads; ++t) {
        storage[t].resize(sizeof(DirectedEnumeration));
        std::memcpy(storage[t].data(), (void*)h->dem, sizeof(DirectedEnumeration));  // bitwise view; never destructed
        clones[t] = reinterpret_cast<DirectedEnumeration*>(storage[t].data());
        lik[t] = new float[n];
        lidx[t] = new int[n];
        clones[t]->likelihoods = lik[t];
        clones[t]->likelihood_indices = lidx[t];
    }
    double secs = run_sharded(nq, nthreads, [&](int t, long lo, long hi) {
        DirectedEnumeration* m = clones[t];
        for (long i = lo; i < hi; ++i) {
            FeaturesVector qv(q + i * (long)d, q + (i + 1) * (long)d);
            ImageInfo qi(-1, -1, qv);
            int best = m->recognize(qi);
            out_idx[i] = best;
            if (out_dist) out_dist[i] = m->bestDistance;
            if (out_below) out_below[i] = m->isFoundLessThreshold ? 1 : 0;
            if (out_evals) out_evals[i] = m->distanceCalcCount;
        }
    });
    for (int t = 1; t < nthreads; ++t) {
        delete[] lik[t];
        delete[] lidx[t];
    }
    return secs;
}

void fir_ref_dem_free(fir_ref_dem* h) {
    if (!h) return;
    delete h->dem;
    if (h->view != &h->db->infos) delete h->view;
    delete h->db;
    delete h;
}

#ifdef FIR_REF_WITH_CLASSIFICATION
// ---- kNN / PNN (classification.cpp:116-226) -------------------------------------------------
// rows: all samples (row-major, n x d, double), labels: class per row.  Fills the file-scope state
// exactly as load_image_dataset (:848-858) leaves it, then runs the verbatim split_train_test (:942-990).
void fir_ref_cls_setup(const double* rows, const int* labels, long n, int d, int n_classes, double fraction, unsigned seed) {
    Silence s;
    fir_ref_set_dim(d);
    dataset.clear();
    indices.clear();
    training_set.clear();
    num_of_cont_features = d;
    num_of_cont_features_orig = d;
    num_of_classes = n_classes;
    for (long i = 0; i < n; ++i) {
        std::vector<FEATURE_TYPE> f(rows + i * (long)d, rows + (i + 1) * (long)d);
        dataset.push_back(Feature_vector(f, labels[i]));
    }
    indices.resize(num_of_classes);
    for (size_t i = 0; i < dataset.size(); ++i) indices[(int)dataset[i].output].push_back(i);
    srand(seed);
    split_train_test(fraction);
}
long fir_ref_cls_counts(int which) {
    if (which == 0) return (long)(dataset.size() - test_set.size());
    return (long)test_set.size();
}
// training rows in the order predict() visits them (class-major, classification.cpp:123-125)
void fir_ref_cls_get_split(long* train_idx, int* train_label, long* test_idx, double* avg) {
    size_t k = 0;
    for (size_t c = 0; c < num_of_classes; ++c)
        for (size_t t = 0; t < training_set[c].size(); ++t) {
            train_idx[k] = (long)training_set[c][t];
            train_label[k] = (int)c;
            ++k;
        }
    for (size_t j = 0; j < test_set.size(); ++j) test_idx[j] = (long)test_set[j];
    for (size_t fi = 0; fi < num_of_cont_features; ++fi) avg[fi] = avgValues[fi];
}
double fir_ref_cls_knn(int K, long first, long count, int nthreads_unused, int* out_label) {
    (void)nthreads_unused;              // file-scope state ⇒ not re-entrant ⇒ serial
    KNNClassifier knn(K);
    auto t1 = std::chrono::high_resolution_clock::now();
    for (long j = 0; j < count; ++j) out_label[j] = knn.predict(tmp_dataset[test_set[first + j]]);
    auto t2 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double>(t2 - t1).count();
}
// labels from the verbatim PNNClassifier::predict (→ predict_bf, :188-226).  The per-class scores are
// locals there (:194), so they are re-evaluated here with the same statements (:189-216) and their
// argmax (:217-225) is checked against the verbatim label; returns -1 on any disagreement.
double fir_ref_cls_pnn(long first, long count, int* out_label, double* out_scores) {
    PNNClassifier pnn(true);
    auto t1 = std::chrono::high_resolution_clock::now();
    for (long j = 0; j < count; ++j) out_label[j] = pnn.predict(tmp_dataset[test_set[first + j]]);
    auto t2 = std::chrono::high_resolution_clock::now();
    if (out_scores) {
        for (long j = 0; j < count; ++j) {
            const Feature_vector& inputFeatures = tmp_dataset[test_set[first + j]];
            size_t total_training_size = dataset.size() - test_set.size();
            double var = 0.00002;
            if (num_of_cont_features > 2000) var /= 10;
            double* outputs = out_scores + j * (long)num_of_classes;
            for (size_t i = 0; i < num_of_classes; ++i) {
                outputs[i] = 0;
                double den = total_training_size;
                for (size_t t = 0; t < training_set[i].size(); ++t) {
                    size_t training_ind = training_set[i][t];
                    FEATURE_TYPE dist = 0;
                    for (size_t fi = 0; fi < num_of_cont_features; ++fi) {
                        FEATURE_TYPE diff = tmp_dataset[training_ind].features[fi] - avgValues[fi];
                        FEATURE_TYPE val = inputFeatures.features[fi] - avgValues[fi];
                        diff -= val;
                        dist += diff * diff;
                    }
                    outputs[i] += exp(-dist / (2 * num_of_cont_features * var));
                }
                outputs[i] /= den;
            }
            double max_output = -DBL_MAX;
            int bestClass = -1;
            for (size_t i = 0; i < num_of_classes; ++i)
                if (max_output < outputs[i]) { max_output = outputs[i]; bestClass = (int)i; }
            if (bestClass != out_label[j]) return -1.0;
        }
    }
    return std::chrono::duration<double>(t2 - t1).count();
}
// PNNClassifier(false): predict() dispatches to predict_sequentional (classification.cpp:228-295, 297-307)
double fir_ref_cls_pnn_seq(long first, long count, int* out_label) {
    PNNClassifier pnn(false);
    auto t1 = std::chrono::high_resolution_clock::now();
    for (long j = 0; j < count; ++j) out_label[j] = pnn.predict(tmp_dataset[test_set[first + j]]);
    auto t2 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double>(t2 - t1).count();
}
void fir_ref_cls_get_std(double* out_std) {
    for (size_t fi = 0; fi < num_of_cont_features; ++fi) out_std[fi] = stdValues[fi];
}
// FPNNClassifier(scale, bf, output_ratio): train() + predict() (classification.cpp:618-791); optionally the trained series
// coefficients a[(fi * C + i) * (2J + 1) ..] and J
int fir_ref_cls_fpnn(double scale, int bf, float output_ratio, long first, long count, int* out_label, double* out_a, int* out_J) {
    Silence s;
    FPNNClassifier f(scale, bf != 0, output_ratio);
    f.train();
    if (out_J) *out_J = (int)f.J;
    if (out_a) std::copy(f.a.begin(), f.a.end(), out_a);
    for (long j = 0; j < count; ++j) out_label[j] = f.predict(tmp_dataset[test_set[first + j]]);
    return (int)f.a.size();
}
// PNNwithClusteringClassifier(no_clusters): train() (per-class k-medoids, :321-388) + predict() (:389-428); the medoids are
// returned as positions in the class-major training order of fir_ref_cls_get_split
int fir_ref_cls_pnn_clustered(int no_clusters, long first, long count, int* out_label, long* out_medoids) {
    Silence s;
    PNNwithClusteringClassifier c(no_clusters);
    c.train();
    int total = 0;
    size_t class_start = 0;
    for (size_t i = 0; i < num_of_classes; ++i) {
        for (size_t m = 0; m < c.clustered_training_set[i].size(); ++m) {
            const size_t row = c.clustered_training_set[i][m];
            size_t pos = 0;
            while (training_set[i][pos] != row) ++pos;
            if (out_medoids) out_medoids[total] = (long)(class_start + pos);
            ++total;
        }
        class_start += training_set[i].size();
    }
    for (long j = 0; j < count; ++j) out_label[j] = c.predict(tmp_dataset[test_set[first + j]]);
    return total;
}
#endif  // FIR_REF_WITH_CLASSIFICATION

}  // extern "C"

// ---- video.cpp (YouTube-Faces experiment): the verbatim loader and driver, compiled from where they lie --------------
#include <unistd.h>
typedef std::map<std::string, std::vector<std::vector<FeaturesVector> > > MapOfVideos;      // video.cpp:22
void loadVideos(MapOfVideos& dbVideos);                                                     // video.cpp:35 (reads VIDEO_FEATURES_FILE in the cwd)
void testYTFRecognition();                                                                  // video.cpp:156

namespace {
struct Cwd {
    char old[4096];
    bool ok;
    explicit Cwd(const char* dir) { ok = getcwd(old, sizeof(old)) != 0 && chdir(dir) == 0; }
    ~Cwd() { if (ok) { int r = chdir(old); (void)r; } }
};
}

extern "C" {

struct fir_ref_videos { MapOfVideos v; int d; };
// loadVideos() run inside `dir`, which must hold the file under the name the reference hard-codes (video.cpp:24-33)
fir_ref_videos* fir_ref_videos_load(const char* dir, int d) {
    Silence s;
    fir_ref_set_dim(d);
    Cwd c(dir);
    if (!c.ok) return 0;
    fir_ref_videos* h = new fir_ref_videos();
    h->d = d;
    loadVideos(h->v);
    return h;
}
void fir_ref_videos_counts(fir_ref_videos* h, long* people, long* videos, long* frames) {
    long nv = 0, nf = 0;
    for (auto& kv : h->v) { nv += (long)kv.second.size(); for (auto& vid : kv.second) nf += (long)vid.size(); }
    *people = (long)h->v.size(); *videos = nv; *frames = nf;
}
// frames flattened in map (name) order; person/video/frame index of each; names joined with '\n'
void fir_ref_videos_get(fir_ref_videos* h, float* frames, int* person, int* video, int* frame, char* names, long names_cap) {
    long r = 0; int p = 0; std::string joined;
    for (auto& kv : h->v) {
        joined += kv.first; joined += '\n';
        for (size_t i = 0; i < kv.second.size(); ++i)
            for (size_t j = 0; j < kv.second[i].size(); ++j) {
                std::memcpy(frames + r * (long)h->d, kv.second[i][j].data(), sizeof(float) * h->d);
                person[r] = p; video[r] = (int)i; frame[r] = (int)j; ++r;
            }
        ++p;
    }
    if (names && names_cap > 0) { std::strncpy(names, joined.c_str(), (size_t)names_cap - 1); names[names_cap - 1] = 0; }
}
void fir_ref_videos_free(fir_ref_videos* h) { delete h; }

// the whole testYTFRecognition() inside `dir` (both hard-coded file names must exist there); what it prints is returned
long fir_ref_ytf_run(const char* dir, int d, char* out, long cap) {
    fir_ref_set_dim(d);
    Cwd c(dir);
    if (!c.ok) return -1;
    std::ostringstream sink;
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
    testYTFRecognition();
    std::cout.rdbuf(old);
    const std::string text = sink.str();
    if (out && cap > 0) { std::strncpy(out, text.c_str(), (size_t)cap - 1); out[cap - 1] = 0; }
    return (long)text.size();
}

}  // extern "C"
