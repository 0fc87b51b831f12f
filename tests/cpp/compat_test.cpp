// Drives the reference-shaped C++ adapters (include/fir_b200_compat.hpp) the way qt_cpp/ann.cpp:24-81 (testANN) and
// classification.cpp:1035-1053 drive the reference classes, and prints machine-readable results for the pytest
// wrapper (tests/test_gpu_compat_cpp.py), which compares them with the oracle.
#include <cstdio>
#include <cstdlib>
#include "fir_b200_compat.hpp"
using namespace fir_compat;

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: compat_test features.txt D pivot0\n"); return 2; }
    fir::features_count() = std::atoi(argv[2]);
    const int pivot0 = std::atoi(argv[3]);
    ImagesDatabase orig, total;
    std::unordered_map<std::string, int> person2index;
    int loaded = loadImages(orig, argv[1], person2index);
    for (size_t i = 0; i < orig.size(); ++i) if (orig[i].size() > 1) total.push_back(orig[i]);      // ann.cpp:34-37
    std::vector<ImageInfo> db, test;
    getTrainingAndTestImages(total, db, test, /*randomize=*/false);
    std::printf("LOADED %d CLASSES %zu DB %zu TEST %zu\n", loaded, total.size(), db.size(), test.size());
    BruteForce bf(db);
    bf.testSetRecognition(test);
    std::vector<int> b = bf.recognize_batch(test);
    std::printf("BF");
    for (size_t i = 0; i < b.size(); ++i) std::printf(" %d", b[i]);
    std::printf("\nBF1 %d RIBF %d RIBF_HALF %d\n", bf.recognize(test[0]), recognize_image_bf(db, test[0]), recognize_image_bf(db, test[0], fir::features_count() / 2));
    std::printf("DIST %.9g\n", test[0].distance(db[3]));
    DirectedEnumeration dem(db, 0.01f, 0, 0, pivot0);
    std::printf("THRESHOLD %.9g\n", dem.threshold());
    dem.setImageCountToCheck((int)(0.2 * db.size()));
    std::vector<int> d = dem.recognize_batch(test);
    std::printf("DEM");
    for (size_t i = 0; i < d.size(); ++i) std::printf(" %d", d[i]);
    std::printf("\n");
    dem.testSetRecognition(test);
    // fp64 kNN / PNN on the same vectors
    std::vector<Feature_vector> rows;
    std::vector<std::vector<size_t> > training_set(total.size());
    for (size_t j = 0; j < db.size(); ++j) {
        rows.push_back(Feature_vector(std::vector<double>(db[j].features.begin(), db[j].features.end()), db[j].classNo));
        training_set[db[j].classNo].push_back(j);
    }
    fir::TrainingSet ts(rows, training_set);
    std::vector<Feature_vector> q;
    for (size_t i = 0; i < test.size(); ++i) q.push_back(Feature_vector(std::vector<double>(test[i].features.begin(), test[i].features.end()), test[i].classNo));
    KNNClassifier knn(3, ts);
    PNNClassifier pnn(ts);
    std::vector<int> k = knn.predict_batch(q), p = pnn.predict_batch(q);
    std::printf("KNN3");
    for (size_t i = 0; i < k.size(); ++i) std::printf(" %d", k[i]);
    std::printf("\nPNN");
    for (size_t i = 0; i < p.size(); ++i) std::printf(" %d", p[i]);
    std::printf("\nPNN1 %d\n", pnn.predict(q[0]));
    PNNClassifier pnn_seq(ts, false);                                    // predict_sequentional
    std::vector<int> ps = pnn_seq.predict_batch(q);
    std::printf("PNNSEQ");
    for (size_t i = 0; i < ps.size(); ++i) std::printf(" %d", ps[i]);
    std::printf("\n");
    // classification.cpp:1002-1011 — the orthogonal-series PNN and the k-medoids-reduced PNN over the same split
    FPNNClassifier fpnn(ts, 1.0, true), fpnn_seq(ts, 0.33, false, 0.9f);
    PNNwithClusteringClassifier pnn_clust(ts, 4);
    fpnn.train(); fpnn_seq.train(); pnn_clust.train();
    std::vector<int> f1 = fpnn.predict_batch(q), f2 = fpnn_seq.predict_batch(q), f3 = pnn_clust.predict_batch(q);
    std::printf("FPNN");
    for (size_t i = 0; i < f1.size(); ++i) std::printf(" %d", f1[i]);
    std::printf("\nFPNNSEQ");
    for (size_t i = 0; i < f2.size(); ++i) std::printf(" %d", f2[i]);
    std::printf("\nPNNCLUST");
    for (size_t i = 0; i < f3.size(); ++i) std::printf(" %d", f3[i]);
    std::printf("\nNAMES %s|%s|%s\n", fpnn.get_name().c_str(), fpnn_seq.get_name().c_str(), pnn_clust.get_name().c_str());
    return 0;
}
