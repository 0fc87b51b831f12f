// fir_b200_compat.hpp — header-only C++ adapters that keep the reference's entry points for the matching
// path and forward them to the C-ABI in fir_b200.h (libfir_b200.so).  A caller of
//   qt_cpp/db_features.h  (FeaturesVector, ImagesDatabase, ImageInfo, feature_distance, loadImages,
//                          getTrainingAndTestImages, recognize_image_bf)
//   qt_cpp/ann.h          (ClassificationMethod, BruteForce, DirectedEnumeration)
//   qt_cpp/classification.cpp's Classifier shape (Feature_vector, Classifier, KNNClassifier, PNNClassifier)
//   qt_cpp/ImageTesting.cpp  (namespace image_testing: Classifier, BruteForceClassifier, ConventionalTWDClassifier,
//                             ProposedTWDClassifier, num_of_unreliable)
// includes this header instead and links -lfir_b200; names, argument meaning and return conventions are the
// reference's (index -1 = no match, matcher objects borrow `std::vector<ImageInfo>&`).  Differences, all additive:
//   * FEATURES_COUNT is a run-time value (fir::features_count()) instead of the macro in qt_cpp/db.h:86,
//     the metric is run time too (fir::metric()) instead of USE_L2_DISTANCE / the `#if` in db_features.cpp:30;
//   * constructors pack the gallery and upload it once; recognize()/predict() are batch-of-1 calls, and
//     testSetRecognition() / recognize_batch() / predict_batch() hand the whole query set to the GPU at once;
//   * the file-scope training state of classification.cpp:53-62 becomes an explicit fir::TrainingSet.
// Everything lives in namespace fir_compat; `using namespace fir_compat;` gives the reference's spelling.
#ifndef FIR_B200_COMPAT_HPP
#define FIR_B200_COMPAT_HPP

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "fir_b200.h"

namespace fir {
inline int& features_count() { static int d = 1536; return d; }          // qt_cpp/db.h:86 default
inline int& metric() { static int m = FIR_L2; return m; }                // qt_cpp/db_features.h:12 default
inline double& fraction() { static double f = 0.03; return f; }          // qt_cpp/db.h:72 (USE_CALTECH)
inline bool& caltech_mode() { static bool c = true; return c; }          // qt_cpp/db.h:11
// GPUs the matchers built from now on spread their gallery over (row shards + NCCL, fir_sharded_*): 1 = one GPU (default),
// 0 = every visible device.  The reference is single-device; this is the one additive knob of the multi-GPU build.
inline int& n_gpus() { static int g = 1; return g; }
inline void check(int status, const char* what) {
    if (status != FIR_OK) throw std::runtime_error(std::string(what) + ": " + fir_last_error_string());
}
}  // namespace fir

namespace fir_compat {

typedef std::vector<float> FeaturesVector;                               // db_features.h:14
typedef std::vector<std::vector<FeaturesVector> > ImagesDatabase;        // db_features.h:15

class ImageInfo;
float feature_distance(const FeaturesVector& lhs, const FeaturesVector& rhs, int start_pos = 0, int end_pos = -1);

class ImageInfo {                                                        // db_features.h:19-29
public:
    ImageInfo(int no, int ind, const FeaturesVector& feat) : classNo(no), indexInDatabase(ind), features(feat) {}
    float distance(const ImageInfo& rhs, int start_pos = 0, int end_pos = -1) const {
        return feature_distance(features, rhs.features, start_pos, end_pos);
    }
    const int classNo, indexInDatabase;
    const FeaturesVector& features;
};

namespace detail {
// a device gallery built from a vector<ImageInfo>; rows are packed in dbImages order (= the tie-break order)
struct PackedGallery {
    fir_gallery* g;
    int d;
    PackedGallery(const std::vector<ImageInfo>& db, int metric, bool build = true) : g(0), d(db.empty() ? 0 : (int)db[0].features.size()) {
        if (db.empty() || !build) return;                                // empty: the reference answers -1 for every query
        std::vector<float> rows((size_t)db.size() * d);
        std::vector<int32_t> labels(db.size());
        for (size_t j = 0; j < db.size(); ++j) {
            std::copy(db[j].features.begin(), db[j].features.begin() + d, rows.begin() + j * (size_t)d);
            labels[j] = db[j].classNo;
        }
        fir::check(fir_gallery_create(rows.data(), labels.data(), (int64_t)db.size(), d, metric, FIR_HOST, 0, &g), "fir_gallery_create");
    }
    ~PackedGallery() { fir_gallery_destroy(g); }
private:
    PackedGallery(const PackedGallery&);
    PackedGallery& operator=(const PackedGallery&);
};
// the same gallery cut into row shards over fir::n_gpus() devices (one process driving them all, csrc/sharded.cu)
struct ShardedPackedGallery {
    fir_sharded* s;
    int d;
    ShardedPackedGallery(const std::vector<ImageInfo>& db, int metric, int n_gpus) : s(0), d(db.empty() ? 0 : (int)db[0].features.size()) {
        if (db.empty() || n_gpus == 1) return;
        std::vector<float> rows((size_t)db.size() * d);
        std::vector<int32_t> labels(db.size());
        for (size_t j = 0; j < db.size(); ++j) {
            std::copy(db[j].features.begin(), db[j].features.begin() + d, rows.begin() + j * (size_t)d);
            labels[j] = db[j].classNo;
        }
        fir::check(fir_sharded_create(rows.data(), labels.data(), (int64_t)db.size(), d, metric, n_gpus, &s), "fir_sharded_create");
    }
    ~ShardedPackedGallery() { fir_sharded_destroy(s); }
private:
    ShardedPackedGallery(const ShardedPackedGallery&);
    ShardedPackedGallery& operator=(const ShardedPackedGallery&);
};
inline std::vector<float> pack_queries(const std::vector<ImageInfo>& q, int d) {
    std::vector<float> rows((size_t)q.size() * d);
    for (size_t i = 0; i < q.size(); ++i) std::copy(q[i].features.begin(), q[i].features.begin() + d, rows.begin() + i * (size_t)d);
    return rows;
}
}  // namespace detail

// feature_distance (db_features.cpp:22-42): a one-row gallery and a one-row query through fir_pair_distances over the
// window [start_pos, end_pos) — the sum starts at start_pos and the mean divides by end_pos - start_pos (:40);
// end_pos < 0 means all dimensions (the reference's default is the FEATURES_COUNT macro).
inline float feature_distance(const FeaturesVector& lhs, const FeaturesVector& rhs, int start_pos, int end_pos) {
    const int d = (int)std::min(lhs.size(), rhs.size());
    if (end_pos < 0 || end_pos > d) end_pos = d;
    if (start_pos < 0 || start_pos >= end_pos) throw std::invalid_argument("feature_distance: empty window");
    fir_gallery* g = 0;
    fir::check(fir_gallery_create(rhs.data() + start_pos, 0, 1, end_pos - start_pos, fir::metric(), FIR_HOST, 0, &g), "fir_gallery_create");
    int32_t idx = 0; float out = 0.f;
    int st = fir_pair_distances(g, lhs.data() + start_pos, 1, &idx, 1, 0, FIR_HOST, &out);
    fir_gallery_destroy(g);
    fir::check(st, "fir_pair_distances");
    return out;
}

// loadImages (db_features.cpp:44-116): same text format (3 lines per image: file, class, D floats), same class
// numbering (first seen), same Caltech background filter; the zeroing + normalisation loop runs on the GPU.
inline int loadImages(ImagesDatabase& imagesDb, std::string features_file, std::unordered_map<std::string, int>& person2indexMap,
                      bool /*early_stop*/ = false) {
    person2indexMap.clear();
    std::ifstream in(features_file.c_str());
    if (!in) return 0;                                                   // missing file ⇒ empty DB, 0 images (db_features.cpp:49,115)
    const int d = fir::features_count();
    std::vector<float> packed;
    std::vector<int> cls;
    std::string file_name, class_name, values;
    while (std::getline(in, file_name) && std::getline(in, class_name) && std::getline(in, values)) {
        size_t first = class_name.find_first_not_of(" \t\n\r\f\v");
        class_name = first == std::string::npos ? std::string() : class_name.substr(first);
        if (fir::caltech_mode() && (class_name.find("BACKGROUND_Google") != std::string::npos || class_name.find("257.clutter") != std::string::npos))
            continue;
        std::unordered_map<std::string, int>::iterator it = person2indexMap.find(class_name);
        if (it == person2indexMap.end()) it = person2indexMap.insert(std::make_pair(class_name, (int)person2indexMap.size())).first;
        std::istringstream ss(values);
        float v = 0.f;
        for (int i = 0; i < d; ++i) { ss >> v; packed.push_back(v); }   // a short line repeats the last parsed value, like the reference's loop
        cls.push_back(it->second);
    }
    const int total = (int)cls.size();
    if (total == 0) return 0;
    fir::check(fir_normalize_rows(packed.data(), total, d, fir::metric(), FIR_HOST, 0), "fir_normalize_rows");
    imagesDb.resize(person2indexMap.size());
    for (int i = 0; i < total; ++i)
        imagesDb[cls[i]].push_back(FeaturesVector(packed.begin() + (size_t)i * d, packed.begin() + (size_t)(i + 1) * d));
    double avg_count = 0;                                                // db_features.cpp:107-112
    for (size_t i = 0; i < imagesDb.size(); ++i) avg_count += imagesDb[i].size();
    avg_count /= imagesDb.size();
    std::cout << "total size=" << imagesDb.size() << " totalImages=" << total << " avg_count=" << avg_count << std::endl;
    return total;
}

// getTrainingAndTestImages (db_features.cpp:117-162): one shuffled index table of 400 shared by every class; the
// first db_size shuffled images of a class go to the gallery, the rest (below 400) to the test set.
inline void getTrainingAndTestImages(const ImagesDatabase& totalImages, std::vector<ImageInfo>& dbImages, std::vector<ImageInfo>& testImages,
                                     bool randomize = true) {
    const int kTable = 400;
    std::vector<int> order(kTable);
    for (int i = 0; i < kTable; ++i) order[i] = i;
    if (randomize) {
#if __cplusplus < 201703L
        std::random_shuffle(order.begin(), order.end());                 // rand()-driven in libstdc++, like the reference
#else
        for (int i = kTable - 1; i > 0; --i) std::swap(order[i], order[std::rand() % (i + 1)]);
#endif
    }
    dbImages.clear();
    testImages.clear();
    int base = 0;
    for (size_t c = 0; c < totalImages.size(); ++c) {
        const int count = (int)totalImages[c].size();
        int db_size = 30;
        if (!fir::caltech_mode()) {
            db_size = (int)std::ceil((float)(count * fir::fraction()));
            if (db_size == count) db_size = count - 1;
            if (db_size == 0) db_size = 1;
        }
        int taken = 0;
        for (int i = 0; i < kTable; ++i) {
            const int j = order[i];
            if (j >= count) continue;
            ImageInfo info((int)c, base + j, totalImages[c][j]);
            if (taken < db_size) dbImages.push_back(info); else testImages.push_back(info);
            ++taken;
        }
        base += count;
    }
}

// recognize_image_bf (db_features.cpp:319-335): 1-NN over the first max_features dimensions (0 = all)
inline int recognize_image_bf(const std::vector<ImageInfo>& dbImages, const ImageInfo& testImageInfo, int max_features = 0) {
    if (dbImages.empty()) return -1;
    detail::PackedGallery pg(dbImages, fir::metric());
    int32_t idx = -1;
    fir::check(fir_search_topk(pg.g, testImageInfo.features.data(), 1, 1, max_features >= pg.d ? 0 : max_features, FIR_PATH_EXACT, FIR_HOST, &idx, 0),
               "fir_search_topk");
    return idx;
}

class ClassificationMethod {                                             // ann.h:9-39
public:
    ClassificationMethod(std::string name, std::vector<ImageInfo>& db) : method_name(name), dbImages(db), avgCheckedPercent(0) {
        imageCountToCheck = (int)dbImages.size();
    }
    virtual ~ClassificationMethod() {}
    virtual int recognize(ImageInfo& testImageInfo) = 0;
    // one GPU call for the whole query set; the default loops recognize()
    virtual std::vector<int> recognize_batch(std::vector<ImageInfo>& testImages) {
        std::vector<int> out(testImages.size());
        for (size_t i = 0; i < testImages.size(); ++i) out[i] = recognize(testImages[i]);
        return out;
    }
    // ann.cpp:94-109: error %, ms per query, checked %
    void testSetRecognition(std::vector<ImageInfo>& testImages) {
        avgCheckedPercent = 0;
        std::chrono::high_resolution_clock::time_point t1 = std::chrono::high_resolution_clock::now();
        std::vector<int> best = recognize_batch(testImages);
        std::chrono::high_resolution_clock::time_point t2 = std::chrono::high_resolution_clock::now();
        int errors = 0;
        for (size_t i = 0; i < testImages.size(); ++i)
            if (best[i] == -1 || testImages[i].classNo != dbImages[best[i]].classNo) ++errors;
        lastErrorRate = testImages.empty() ? 0.0 : 100.0 * errors / testImages.size();
        const double ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
        std::cout << method_name << " error=" << lastErrorRate << "% total_time (ms)" << (testImages.empty() ? 0.0 : ms / testImages.size())
                  << " checkedPercent=" << (avgCheckedPercent > 0 ? avgCheckedPercent / testImages.size() : -1) << std::endl;
    }
    virtual void setImageCountToCheck(int count) {                       // ann.h:20-22
        imageCountToCheck = (count > 0 && count < (int)dbImages.size()) ? count : (int)dbImages.size();
    }
    static float getThreshold(std::vector<float>& otherClassesDists, float falseAcceptRate) {   // ann.cpp:84-93
        int ind = (int)(otherClassesDists.size() * falseAcceptRate);
        std::nth_element(otherClassesDists.begin(), otherClassesDists.begin() + ind, otherClassesDists.end());
        return otherClassesDists[ind];
    }
    double lastErrorRate = 0;
protected:
    std::string method_name;
    std::vector<ImageInfo>& dbImages;
    int distanceCalcCount = 0;
    float avgCheckedPercent;
    int imageCountToCheck;
};

class BruteForce : public ClassificationMethod {                         // ann.h:42-47, ann.cpp:113-126
public:
    // fir::n_gpus() == 1: one device gallery; otherwise row shards over the GPUs of the box + NCCL candidate merge
    BruteForce(std::vector<ImageInfo>& db) : ClassificationMethod("BF", db), sg(db, fir::metric(), fir::n_gpus()),
                                             pg(db, fir::metric(), sg.s == 0) {}
    int recognize(ImageInfo& testImage) {
        std::vector<ImageInfo> one(1, testImage);
        return recognize_batch(one)[0];
    }
    std::vector<int> recognize_batch(std::vector<ImageInfo>& testImages) {
        std::vector<int> out(testImages.size(), -1);
        if ((!pg.g && !sg.s) || testImages.empty()) return out;
        std::vector<float> q = detail::pack_queries(testImages, sg.s ? sg.d : pg.d);
        std::vector<int32_t> idx(testImages.size());
        if (sg.s) fir::check(fir_sharded_search_topk(sg.s, q.data(), (int64_t)testImages.size(), 1, FIR_PATH_AUTO, idx.data(), 0), "fir_sharded_search_topk");
        else fir::check(fir_search_topk(pg.g, q.data(), (int64_t)testImages.size(), 1, 0, FIR_PATH_AUTO, FIR_HOST, idx.data(), 0), "fir_search_topk");
        distanceCalcCount = (int)dbImages.size();
        for (size_t i = 0; i < out.size(); ++i) out[i] = idx[i];
        return out;
    }
private:
    detail::ShardedPackedGallery sg;
    detail::PackedGallery pg;
};

class DirectedEnumeration : public ClassificationMethod {                // ann.h:61-100, ann.cpp:270-507
public:
    // The reference's first pivot is the head of std::random_shuffle over 0..N-1 (ann.cpp:366-369; the rest of that
    // shuffle is overwritten by the farthest-point chain, :328-330).  With the reference's four arguments the same shuffle
    // is drawn here — same rand() consumption, so under the same srand() the pivot chain is the reference's (libstdc++).
    // pivot0 >= 0 fixes the first pivot instead; pivot0 < 0 with seed != 0 derives it from `seed` without touching rand().
    DirectedEnumeration(std::vector<ImageInfo>& faceImages, float falseAcceptRate = 0.01f, float threshold = 0, int imageCountToCheck = 0,
                        int pivot0 = -1, unsigned seed = 0)
        : ClassificationMethod("dem", faceImages), isFoundLessThreshold(false), bestDistance(0), sg(faceImages, fir::metric(), fir::n_gpus()),
          pg(faceImages, fir::metric(), sg.s == 0), dem(0), sdem(0) {
        setImageCountToCheck(imageCountToCheck);
        if (pivot0 < 0 && seed == 0 && !faceImages.empty()) {
            std::vector<int> indices(faceImages.size());
            for (size_t i = 0; i < indices.size(); ++i) indices[i] = (int)i;
#if __cplusplus < 201703L
            std::random_shuffle(indices.begin(), indices.end());
#else
            for (size_t i = indices.size() - 1; i > 0; --i) std::swap(indices[i], indices[std::rand() % (i + 1)]);
#endif
            pivot0 = indices[0];
        }
        fir_dem_params p;
        p.pivot0 = pivot0; p.seed = seed; p.false_accept_rate = falseAcceptRate; p.threshold = threshold; p.max_chain = 0; p.max_pivots = 0;
        std::chrono::high_resolution_clock::time_point t1 = std::chrono::high_resolution_clock::now();
        if (sg.s) fir::check(fir_sharded_dem_build(sg.s, &p, &sdem), "fir_sharded_dem_build");      // fir::n_gpus() != 1: ONE index over the row shards
        else fir::check(fir_dem_build(pg.g, &p, &dem), "fir_dem_build");
        std::cout << "init took " << std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - t1).count()
                  << " milliseconds" << std::endl;
    }
    ~DirectedEnumeration() { fir_dem_destroy(dem); fir_sharded_dem_destroy(sdem); }
    int recognize(ImageInfo& testImage) {
        std::vector<ImageInfo> one(1, testImage);
        return recognize_batch(one)[0];
    }
    std::vector<int> recognize_batch(std::vector<ImageInfo>& testImages) {
        const size_t nq = testImages.size();
        std::vector<int> out(nq, -1);
        if (nq == 0) return out;
        std::vector<float> q = detail::pack_queries(testImages, sg.s ? sg.d : pg.d);
        std::vector<int32_t> idx(nq), evals(nq);
        std::vector<float> dist(nq);
        std::vector<uint8_t> below(nq);
        if (sdem) fir::check(fir_sharded_dem_search(sdem, q.data(), (int64_t)nq, imageCountToCheck, idx.data(), dist.data(), below.data(), evals.data()), "fir_sharded_dem_search");
        else fir::check(fir_dem_search(dem, q.data(), (int64_t)nq, imageCountToCheck, FIR_HOST, idx.data(), dist.data(), below.data(), evals.data()), "fir_dem_search");
        for (size_t i = 0; i < nq; ++i) {
            out[i] = idx[i];
            avgCheckedPercent += (float)(100. * evals[i] / dbImages.size());        // ann.cpp:505
        }
        isFoundLessThreshold = below[nq - 1] != 0; bestDistance = dist[nq - 1]; distanceCalcCount = evals[nq - 1];
        return out;
    }
    float threshold() const { float t = 0; if (sdem) fir_sharded_dem_info(sdem, 0, 0, &t); else fir_dem_info(dem, 0, 0, &t); return t; }
    bool isFoundLessThreshold;
    float bestDistance;
private:
    detail::ShardedPackedGallery sg;
    detail::PackedGallery pg;
    fir_dem* dem;
    fir_sharded_dem* sdem;
};

// ---- classification.cpp ------------------------------------------------------------------------------
class Feature_vector {                                                   // classification.cpp:35-43
public:
    Feature_vector(const std::vector<double>& fv, double out) : features(fv), output(out) {}
    std::vector<double> features;
    double output;
};
}  // namespace fir_compat

namespace fir {
// the state classification.cpp keeps in file-scope globals (:53-62) after split_train_test (:942-990):
// training rows in class-major order, their labels, the per-feature mean of the training rows
struct TrainingSet {
    fir_classifier* c;
    int n_classes, d;
    std::vector<double> packed, avg, stddev;                             // raw training rows (class-major), avgValues, stdValues
    std::vector<int32_t> labels;
    TrainingSet(const std::vector<fir_compat::Feature_vector>& rows, const std::vector<std::vector<size_t> >& training_set) : c(0) {
        n_classes = (int)training_set.size();
        d = rows.empty() ? 0 : (int)rows[0].features.size();
        avg.assign((size_t)d, 0.0);
        stddev.assign((size_t)d, 0.0);
        size_t count = 0;
        for (size_t k = 0; k < training_set.size(); ++k)
            for (size_t t = 0; t < training_set[k].size(); ++t) {
                const std::vector<double>& f = rows[training_set[k][t]].features;
                packed.insert(packed.end(), f.begin(), f.begin() + d);
                labels.push_back((int32_t)k);
                ++count;
            }
        for (int fi = 0; fi < d; ++fi) {                                 // avgValues, classification.cpp:969-987 (same summation order)
            double s = 0, s2 = 0;
            for (size_t r = 0; r < count; ++r) { const double f = packed[r * (size_t)d + fi]; s += f; s2 += f * f; }
            avg[fi] = s / (int)count;
            stddev[fi] = std::sqrt((s2 - avg[fi] * avg[fi] * (int)count) / ((int)count - 1));      // stdValues, :988
        }
        check(fir_classifier_create(packed.data(), labels.data(), (int64_t)count, d, n_classes, avg.data(), &c), "fir_classifier_create");
    }
    ~TrainingSet() { fir_classifier_destroy(c); }
private:
    TrainingSet(const TrainingSet&);
    TrainingSet& operator=(const TrainingSet&);
};
}  // namespace fir

namespace fir_compat {
class Classifier {                                                       // classification.cpp:82-95
public:
    Classifier(std::string name, fir::TrainingSet& ts) : method_name(name), train_set(ts) {}
    virtual ~Classifier() {}
    virtual void train() {}
    virtual int predict(const Feature_vector& inputFeatures) {
        std::vector<Feature_vector> one(1, inputFeatures);
        return predict_batch(one)[0];
    }
    virtual std::vector<int> predict_batch(const std::vector<Feature_vector>& inputs) = 0;
    std::string get_name() { return method_name; }
protected:
    std::string method_name;
    fir::TrainingSet& train_set;
    std::vector<double> pack(const std::vector<Feature_vector>& in) const {
        std::vector<double> q;
        for (size_t i = 0; i < in.size(); ++i) q.insert(q.end(), in[i].features.begin(), in[i].features.begin() + train_set.d);
        return q;
    }
};

class KNNClassifier : public Classifier {                                // classification.cpp:108-170
public:
    KNNClassifier(int k, fir::TrainingSet& ts) : Classifier("k-NN, " + std::to_string(k), ts), K(k) {}
    std::vector<int> predict_batch(const std::vector<Feature_vector>& inputs) {
        std::vector<double> q = pack(inputs);
        std::vector<int32_t> lab(inputs.size());
        fir::check(fir_classifier_knn(train_set.c, q.data(), (int64_t)inputs.size(), K, lab.data()), "fir_classifier_knn");
        return std::vector<int>(lab.begin(), lab.end());
    }
private:
    int K;
};

class PNNClassifier : public Classifier {                                // classification.cpp:173-226 (predict_bf)
public:
    PNNClassifier(fir::TrainingSet& ts, bool bf = true, std::string name = "PNN") : Classifier(name + (bf ? "" : " (seq)"), ts), bruteforce(bf) {}
    std::vector<int> predict_batch(const std::vector<Feature_vector>& inputs) {
        std::vector<double> q = pack(inputs);
        std::vector<int32_t> lab(inputs.size());
        if (!bruteforce) {                                               // predict_sequentional, classification.cpp:228-295
            fir::check(fir_classifier_pnn_sequential(train_set.c, q.data(), (int64_t)inputs.size(), lab.data()), "fir_classifier_pnn_sequential");
            return std::vector<int>(lab.begin(), lab.end());
        }
        scores.assign(inputs.size() * (size_t)train_set.n_classes, 0.0);
        fir::check(fir_classifier_pnn(train_set.c, q.data(), (int64_t)inputs.size(), scores.data(), lab.data()), "fir_classifier_pnn");
        return std::vector<int>(lab.begin(), lab.end());
    }
    std::vector<double> scores;                                          // per-class outputs of the last call (locals in the reference, :194)
private:
    bool bruteforce;
};

class PNNwithClusteringClassifier : public Classifier {                  // classification.cpp:311-428
public:
    PNNwithClusteringClassifier(fir::TrainingSet& ts, int no_clusters) : Classifier(label("PNN with clustering", no_clusters), ts), num_clusters(no_clusters), reduced(0) {}
    ~PNNwithClusteringClassifier() { fir_classifier_destroy(reduced); }
    void train() {                                                       // per-class k-medoids (:321-388), then the reduced Parzen set
        const int64_t n = (int64_t)train_set.labels.size();
        std::vector<int64_t> sel((size_t)n);
        int64_t cnt = 0;
        fir::check(fir_kmedoids_select(train_set.packed.data(), train_set.labels.data(), n, train_set.d, train_set.n_classes, num_clusters, sel.data(), &cnt),
                   "fir_kmedoids_select");
        std::vector<double> rows;
        std::vector<int32_t> lab;
        for (int64_t i = 0; i < cnt; ++i) {
            rows.insert(rows.end(), train_set.packed.begin() + sel[i] * train_set.d, train_set.packed.begin() + (sel[i] + 1) * train_set.d);
            lab.push_back(train_set.labels[sel[i]]);
        }
        fir_classifier_destroy(reduced); reduced = 0;
        fir::check(fir_classifier_create(rows.data(), lab.data(), cnt, train_set.d, train_set.n_classes, train_set.avg.data(), &reduced), "fir_classifier_create");
        fir::check(fir_classifier_set_total(reduced, n), "fir_classifier_set_total");      // den = total_training_size (:393)
    }
    std::vector<int> predict_batch(const std::vector<Feature_vector>& inputs) {
        if (!reduced) train();
        std::vector<double> q = pack(inputs);
        std::vector<int32_t> lab(inputs.size());
        fir::check(fir_classifier_pnn(reduced, q.data(), (int64_t)inputs.size(), 0, lab.data()), "fir_classifier_pnn");
        return std::vector<int>(lab.begin(), lab.end());
    }
private:
    static std::string label(const char* prefix, int v) { std::ostringstream os; os << prefix << ", " << v; return os.str(); }
    int num_clusters;
    fir_classifier* reduced;
};

class FPNNClassifier : public Classifier {                               // classification.cpp:618-791
public:
    FPNNClassifier(fir::TrainingSet& ts, double scale = 1.0, bool bf = true, float output_ratio = 0.9f)
        : Classifier(label(scale) + (bf ? "" : " (seq)"), ts), features_scale(scale), bruteforce(bf), ratio(output_ratio), f(0) {}
    ~FPNNClassifier() { fir_fpnn_destroy(f); }
    void train() {                                                       // :658-695
        fir_fpnn_destroy(f); f = 0;
        fir::check(fir_fpnn_create(train_set.packed.data(), train_set.labels.data(), (int64_t)train_set.labels.size(), train_set.d, train_set.n_classes,
                                   train_set.avg.data(), train_set.stddev.data(), features_scale, &f), "fir_fpnn_create");
    }
    std::vector<int> predict_batch(const std::vector<Feature_vector>& inputs) {
        if (!f) train();
        std::vector<double> q = pack(inputs);
        std::vector<int32_t> lab(inputs.size());
        fir::check(fir_fpnn_predict(f, q.data(), (int64_t)inputs.size(), bruteforce ? 0 : 1, ratio, lab.data()), "fir_fpnn_predict");
        return std::vector<int>(lab.begin(), lab.end());
    }
private:
    static std::string label(double v) { std::ostringstream os; os << "FPNN" << ", " << v; return os.str(); }
    double features_scale;
    bool bruteforce;
    float ratio;
    fir_fpnn* f;
};

// ---- qt_cpp/video.cpp: the YouTube-Faces experiment, a second caller of the same matching path ------------------
typedef std::map<std::string, std::vector<std::vector<FeaturesVector> > > MapOfVideos;       // video.cpp:22

// loadVideos (video.cpp:35-95): per person a name line and a video count; per video a frame count; per frame a file-name
// line and D floats.  The zeroing + normalisation of every frame (:64-84) runs on the GPU in one call.
inline void loadVideos(MapOfVideos& dbVideos, const std::string& video_features_file) {
    std::ifstream ifs(video_features_file.c_str());
    if (!ifs) return;
    const int d = fir::features_count();
    std::vector<float> packed;                                           // every frame of the file, in file order
    struct Slot { std::string person; int video, frame; };
    std::vector<Slot> slots;
    int total_videos = 0;
    while (ifs) {
        std::string fileName, personName, feat_str;
        if (!std::getline(ifs, personName)) break;
        personName.erase(0, personName.find_first_not_of(" \t\n\r\f\v\r\n"));
        int videos_count = 0;
        ifs >> videos_count;
        dbVideos.insert(std::make_pair(personName, std::vector<std::vector<FeaturesVector> >()));
        std::vector<std::vector<FeaturesVector> >& person_videos = dbVideos[personName];
        person_videos.resize(videos_count);
        for (int i = 0; i < videos_count; ++i) {
            int frames_count = 0;
            ifs >> frames_count;
            person_videos[i].resize(frames_count);
            if (!std::getline(ifs, fileName)) break;                     // rest of the count line
            for (int j = 0; j < frames_count; ++j) {
                if (!std::getline(ifs, fileName)) break;
                if (!std::getline(ifs, feat_str)) break;
                std::istringstream iss(feat_str);
                float v = 0.f;
                for (int k = 0; k < d; ++k) { iss >> v; packed.push_back(v); }
                Slot s; s.person = personName; s.video = i; s.frame = j;
                slots.push_back(s);
            }
        }
        total_videos += videos_count;
    }
    if (!slots.empty())
        fir::check(fir_normalize_rows(packed.data(), (int64_t)slots.size(), d, fir::metric() == FIR_L2 ? (int)FIR_L2 : (int)FIR_NORM_VIDEO_SUMSQ, FIR_HOST, 0),
                   "fir_normalize_rows");
    for (size_t i = 0; i < slots.size(); ++i)
        dbVideos[slots[i].person][slots[i].video][slots[i].frame].assign(packed.begin() + i * (size_t)d, packed.begin() + (i + 1) * (size_t)d);
    std::cout << "total size=" << dbVideos.size() << " totalVideos=" << total_videos << " totalImages=" << slots.size() << std::endl;
}

// The split of testYTFRecognition (video.cpp:167-236): people present in both sets, every 10th frame of every video as a
// query, the still images as the gallery (in the iteration order of the surviving person2indexMap, like the reference).
inline void buildYTFSplit(ImagesDatabase& totalImages, std::unordered_map<std::string, int>& person2indexMap, MapOfVideos& videos,
                          std::vector<ImageInfo>& dbImages, std::vector<ImageInfo>& testImages) {
    std::vector<std::string> dbNames, videoNames;
    for (std::unordered_map<std::string, int>::iterator it = person2indexMap.begin(); it != person2indexMap.end(); ++it) dbNames.push_back(it->first);
    for (MapOfVideos::iterator it = videos.begin(); it != videos.end(); ++it) videoNames.push_back(it->first);
    std::sort(videoNames.begin(), videoNames.end());
    std::sort(dbNames.begin(), dbNames.end());
    std::vector<std::string> commonNames(videos.size());
    commonNames.resize(std::set_intersection(videoNames.begin(), videoNames.end(), dbNames.begin(), dbNames.end(), commonNames.begin()) - commonNames.begin());
    std::cout << "lfw names size=" << dbNames.size() << " YTF names size=" << videoNames.size() << " common names size=" << commonNames.size() << std::endl;
    std::vector<std::string> listToRemove;
    std::set_symmetric_difference(videoNames.begin(), videoNames.end(), dbNames.begin(), dbNames.end(), std::back_inserter(listToRemove));
    for (size_t i = 0; i < listToRemove.size(); ++i) { person2indexMap.erase(listToRemove[i]); videos.erase(listToRemove[i]); }
    std::unordered_map<std::string, int> person2indexMapNew;
    int class_index = 0;
    for (MapOfVideos::iterator iter = videos.begin(); iter != videos.end(); ++iter) {
        person2indexMapNew.insert(std::make_pair(iter->first, class_index));
        for (size_t v = 0; v < iter->second.size(); ++v)
            for (int ind = 0; ind < (int)iter->second[v].size(); ind += 10)            // :216
                testImages.push_back(ImageInfo(class_index, ind, iter->second[v][ind]));
        ++class_index;
    }
    int lfw_size = 0;
    for (std::unordered_map<std::string, int>::iterator pi = person2indexMap.begin(); pi != person2indexMap.end(); ++pi) {
        const int new_class_ind = person2indexMapNew[pi->first];
        const int class_ind = pi->second;
        for (int ind = 0; ind < (int)totalImages[class_ind].size(); ++ind) dbImages.push_back(ImageInfo(new_class_ind, ind, totalImages[class_ind][ind]));
        ++lfw_size;
    }
    std::cout << "lfw names size=" << lfw_size << " YTF names size=" << videos.size() << " removed=" << listToRemove.size() << std::endl;
    std::cout << "dbSize=" << dbImages.size() << " testSize=" << testImages.size() << std::endl;
}

// ---- qt_cpp/ImageTesting.cpp: Classifier (:35-48) and its three matching classifiers ----------------------------
// ImageTesting.cpp has its own `class Classifier` (train(&dbImages) / recognize(testImageInfo) → class); it lives in a nested
// namespace here because classification.cpp's Classifier above has the same name.  train() packs and uploads the gallery;
// recognize() is a batch-of-1 call and recognize_batch() hands the whole test set to the GPU.  num_of_unreliable (:33) is
// the file-scope counter testRecognitionMethod resets and prints (:458,:473).
namespace image_testing {

inline int& num_of_unreliable() { static int n = 0; return n; }

class Classifier {
public:
    Classifier(std::string n) : pDbImages(0), pg(0), name(n) {}
    virtual ~Classifier() { delete pg; }
    virtual void train(std::vector<ImageInfo>* pDb) {
        pDbImages = pDb;
        delete pg;
        pg = new detail::PackedGallery(*pDb, fir::metric());
    }
    virtual int recognize(ImageInfo& testImageInfo) {
        std::vector<ImageInfo> one(1, testImageInfo);
        return recognize_batch(one)[0];
    }
    virtual std::vector<int> recognize_batch(std::vector<ImageInfo>& testImages) = 0;
    std::string get_name() { return name; }
protected:
    std::vector<ImageInfo>* pDbImages;
    detail::PackedGallery* pg;
    static std::string build_name(std::string prefix, int param) {
        std::ostringstream os;
        os << prefix << ", " << param;
        return os.str();
    }
    // run one of the C-ABI classifiers over a batch; counts the unreliable ones like the reference's global does
    template <typename Call> std::vector<int> run(std::vector<ImageInfo>& testImages, Call call) {
        std::vector<int> out(testImages.size(), -1);
        if (!pg || !pg->g || testImages.empty()) return out;
        std::vector<float> q = detail::pack_queries(testImages, pg->d);
        std::vector<int32_t> lab(testImages.size());
        std::vector<uint8_t> unrel(testImages.size());
        call(q.data(), (int64_t)testImages.size(), lab.data(), unrel.data());
        for (size_t i = 0; i < out.size(); ++i) { out[i] = lab[i]; num_of_unreliable() += unrel[i]; }
        return out;
    }
private:
    std::string name;
    Classifier(const Classifier&);
    Classifier& operator=(const Classifier&);
};

class BruteForceClassifier : public Classifier {                         // ImageTesting.cpp:58-71
public:
    BruteForceClassifier(int max_feats = fir::features_count()) : Classifier(Classifier::build_name("BF", max_feats)), max_features(max_feats) {}
    std::vector<int> recognize_batch(std::vector<ImageInfo>& testImages) {
        std::vector<int> out(testImages.size(), -1);
        if (!pg || !pg->g || testImages.empty()) return out;
        std::vector<float> q = detail::pack_queries(testImages, pg->d);
        std::vector<int32_t> idx(testImages.size());
        // recognize_image_bf treats max_features == 0 as "all" and divides by the prefix length otherwise (db_features.cpp:319-335)
        fir::check(fir_search_topk(pg->g, q.data(), (int64_t)testImages.size(), 1, max_features >= pg->d ? 0 : max_features, FIR_PATH_AUTO,
                                   FIR_HOST, idx.data(), 0), "fir_search_topk");
        for (size_t i = 0; i < out.size(); ++i) out[i] = idx[i] >= 0 ? (*pDbImages)[idx[i]].classNo : -1;
        return out;
    }
private:
    int max_features;
};

class ConventionalTWDClassifier : public Classifier {                    // ImageTesting.cpp:74-186
public:
    enum class TWD_Type { Posteriors, DistDiff, DistRatio };
    ConventionalTWDClassifier(int cls_num, TWD_Type t, double th, int feat_count = 64)
        : Classifier(build_name(t, th)), num_of_classes(cls_num), reduced_features_count(feat_count), threshold(th), type(t) {}
    void train(std::vector<ImageInfo>* pDb) {
        Classifier::train(pDb);
        if (pg->g) fir::check(fir_gallery_set_num_classes(pg->g, num_of_classes), "fir_gallery_set_num_classes");   // vector<double> probabs(num_of_classes), :114
    }
    std::vector<int> recognize_batch(std::vector<ImageInfo>& testImages) {
        const int t = type == TWD_Type::Posteriors ? FIR_TWD_POSTERIORS : type == TWD_Type::DistDiff ? FIR_TWD_DIST_DIFF : FIR_TWD_DIST_RATIO;
        fir_gallery* g = pg ? pg->g : 0;
        const double th = threshold; const int fc = reduced_features_count;
        return run(testImages, [=](const float* q, int64_t nq, int32_t* lab, uint8_t* unrel) {
            fir::check(fir_twd_conventional(g, q, nq, t, th, fc, 256, FIR_HOST, 0, lab, unrel), "fir_twd_conventional");
        });
    }
private:
    int num_of_classes, reduced_features_count;
    double threshold;
    TWD_Type type;
    static std::string build_name(TWD_Type type, double threshold) {
        std::ostringstream os;
        os << (type == TWD_Type::Posteriors ? "TWD posteriors" : type == TWD_Type::DistDiff ? "TWD diff" : "TWD ratio") << ", " << threshold;
        return os.str();
    }
};

class ProposedTWDClassifier : public Classifier {                        // ImageTesting.cpp:188-288
public:
    ProposedTWDClassifier(int cls_num, int feat_count, double th)
        : Classifier(build_name(feat_count, th)), num_of_classes(cls_num), reduced_features_count(feat_count), threshold(th) {}
    std::vector<int> recognize_batch(std::vector<ImageInfo>& testImages) {
        fir_gallery* g = pg ? pg->g : 0;
        const double th = threshold; const int fc = reduced_features_count;
        return run(testImages, [=](const float* q, int64_t nq, int32_t* lab, uint8_t* unrel) {
            fir::check(fir_twd_proposed(g, q, nq, fc, th, 256, FIR_HOST, 0, lab, unrel), "fir_twd_proposed");
        });
    }
private:
    int num_of_classes, reduced_features_count;
    double threshold;                                                    // the reference stores 1/th (:191); the C-ABI takes th
    static std::string build_name(int feat_count, double threshold) {
        std::ostringstream os;
        os << "Proposed TWD, " << feat_count << ", " << threshold;
        return os.str();
    }
};

}  // namespace image_testing

}  // namespace fir_compat

#endif  // FIR_B200_COMPAT_HPP
