// BASELINE config 1 at the reference's own shape, driven like testANN (qt_cpp/ann.cpp:24-81): the 3-line text features
// file → loadImages → getTrainingAndTestImages(randomize = true) under srand(13) → BruteForce → DirectedEnumeration.
// Prints the split, the per-query answers and the error rates for tests/test_gpu_c1_shape.py, which compares them with
// the unmodified reference (oracle/_ref) run on the same file with the same seeds.
#include <cstdio>
#include <cstdlib>
#include "fir_b200_compat.hpp"
using namespace fir_compat;

int main(int argc, char** argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: c1_test features.txt D split_seed dem_seed [n_gpus]\n"); return 2; }
    fir::features_count() = std::atoi(argv[2]);
    if (argc > 5) fir::n_gpus() = std::atoi(argv[5]);
    ImagesDatabase orig, total;
    std::unordered_map<std::string, int> person2index;
    const int loaded = loadImages(orig, argv[1], person2index);
    for (size_t i = 0; i < orig.size(); ++i) if (orig[i].size() > 1) total.push_back(orig[i]);      // ann.cpp:34-37
    std::vector<ImageInfo> db, test;
    srand(std::atoi(argv[3]));
    getTrainingAndTestImages(total, db, test);                                                     // randomize = true, the reference's default
    std::printf("LOADED %d CLASSES %zu DB %zu TEST %zu\n", loaded, total.size(), db.size(), test.size());
    std::printf("DBIDX");
    for (size_t i = 0; i < db.size(); ++i) std::printf(" %d", db[i].indexInDatabase);
    std::printf("\nTESTIDX");
    for (size_t i = 0; i < test.size(); ++i) std::printf(" %d", test[i].indexInDatabase);
    std::printf("\n");
    BruteForce bf(db);
    bf.testSetRecognition(test);
    std::vector<int> b = bf.recognize_batch(test);
    std::printf("BF");
    for (size_t i = 0; i < b.size(); ++i) std::printf(" %d", b[i]);
    std::printf("\nBFERR %.9g\n", bf.lastErrorRate);
    std::printf("WINDOW %.9g %.9g\n", feature_distance(test[0].features, db[1].features, 64, 320), test[1].distance(db[2], 100, 101));
    srand(std::atoi(argv[4]));
    DirectedEnumeration dem(db);                                                                   // the reference's default arguments
    std::printf("THRESHOLD %.9g\n", dem.threshold());
    const double ratios[] = {0.05, 0.2};
    for (int r = 0; r < 2; ++r) {
        dem.setImageCountToCheck((int)(ratios[r] * db.size()));
        std::vector<int> d = dem.recognize_batch(test);
        std::printf("DEM%d", r);
        for (size_t i = 0; i < d.size(); ++i) std::printf(" %d", d[i]);
        std::printf("\n");
        dem.testSetRecognition(test);
        std::printf("DEMERR%d %.9g\n", r, dem.lastErrorRate);
    }
    return 0;
}
