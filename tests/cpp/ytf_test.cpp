// Drives the video.cpp-shaped adapters (loadVideos, buildYTFSplit) and the matchers the way testYTFRecognition
// (qt_cpp/video.cpp:156-267) does.  Prints machine-readable lines for tests/test_gpu_compat_cpp.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "fir_b200_compat.hpp"
using namespace fir_compat;

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: ytf_test dir D pivot0\n"); return 2; }
    const std::string dir = argv[1];
    fir::features_count() = std::atoi(argv[2]);
    const int pivot0 = std::atoi(argv[3]);
    ImagesDatabase totalImages;
    std::unordered_map<std::string, int> person2indexMap;
    loadImages(totalImages, dir + "/dnn_vgg_features_all_mean.txt", person2indexMap);        // TRAIN_FEATURES_FILE, video.cpp:32
    MapOfVideos videos;
    loadVideos(videos, dir + "/vgg_mean_dnn_features.txt");                                  // VIDEO_FEATURES_FILE, video.cpp:30
    std::vector<ImageInfo> dbImages, testImages;
    buildYTFSplit(totalImages, person2indexMap, videos, dbImages, testImages);
    unsigned long long sum = 0;
    for (size_t i = 0; i < testImages.size(); ++i)
        for (size_t k = 0; k < testImages[i].features.size(); ++k) { unsigned u; std::memcpy(&u, &testImages[i].features[k], 4); sum += u; }
    std::printf("TESTBITS %llu\n", sum);
    BruteForce bf(dbImages);
    bf.testSetRecognition(testImages);
    std::vector<int> b = bf.recognize_batch(testImages);
    std::printf("BFCLASS");
    for (size_t i = 0; i < b.size(); ++i) std::printf(" %d", b[i] < 0 ? -1 : dbImages[b[i]].classNo);
    std::printf("\nTESTCLASS");
    for (size_t i = 0; i < testImages.size(); ++i) std::printf(" %d", testImages[i].classNo);
    std::printf("\n");
    DirectedEnumeration dem(dbImages, 0.01f, 0, 0, pivot0);
    for (double ratio = 0.1; ratio <= 0.7; ratio += 0.1) {                                   // video.cpp:253
        dem.setImageCountToCheck((int)(ratio * dbImages.size()));
        std::cout << "ratio" << ratio << std::endl;
        dem.testSetRecognition(testImages);
    }
    std::vector<int> dd = dem.recognize_batch(testImages);
    std::printf("DEMCLASS");
    for (size_t i = 0; i < dd.size(); ++i) std::printf(" %d", dd[i] < 0 ? -1 : dbImages[dd[i]].classNo);
    std::printf("\n");
    return 0;
}
