// Drives the ImageTesting.cpp-shaped adapters (fir_compat::image_testing) the way testRecognition / testRecognitionMethod
// (qt_cpp/ImageTesting.cpp:439-548) drive the reference classes: one classifier list, train(&dbImages), recognize per test
// image, num_of_unreliable read afterwards.  Prints machine-readable lines for tests/test_gpu_compat_cpp.py.
#include <cstdio>
#include <cstdlib>
#include "fir_b200_compat.hpp"
using namespace fir_compat;
using namespace fir_compat::image_testing;

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: twd_test features.txt D\n"); return 2; }
    fir::features_count() = std::atoi(argv[2]);
    ImagesDatabase totalImages;
    std::unordered_map<std::string, int> person2indexMap;
    loadImages(totalImages, argv[1], person2indexMap);
    const int num_of_classes = (int)totalImages.size();
    std::vector<ImageInfo> dbImages, testImages;
    getTrainingAndTestImages(totalImages, dbImages, testImages, /*randomize=*/false);
    std::vector<image_testing::Classifier*> classifiers;                 // the list of ImageTesting.cpp:525-534
    classifiers.push_back(new BruteForceClassifier());
    classifiers.push_back(new BruteForceClassifier(64));
    classifiers.push_back(new BruteForceClassifier(256));
    classifiers.push_back(new ConventionalTWDClassifier(num_of_classes, ConventionalTWDClassifier::TWD_Type::Posteriors, 0.24));
    classifiers.push_back(new ConventionalTWDClassifier(num_of_classes, ConventionalTWDClassifier::TWD_Type::DistDiff, 0.003));
    classifiers.push_back(new ConventionalTWDClassifier(num_of_classes, ConventionalTWDClassifier::TWD_Type::DistRatio, 0.7));
    classifiers.push_back(new ProposedTWDClassifier(num_of_classes, 32, 0.7));
    classifiers.push_back(new ProposedTWDClassifier(num_of_classes, 64, 0.7));
    std::printf("SIZES %d %zu %zu\n", num_of_classes, dbImages.size(), testImages.size());
    for (size_t c = 0; c < classifiers.size(); ++c) {
        image_testing::Classifier* cl = classifiers[c];
        cl->train(&dbImages);
        num_of_unreliable() = 0;
        std::vector<int> batch = cl->recognize_batch(testImages);
        const int unreliable_batch = num_of_unreliable();
        std::printf("NAME%zu %s\nCLS%zu", c, cl->get_name().c_str(), c);
        for (size_t i = 0; i < batch.size(); ++i) std::printf(" %d", batch[i]);
        num_of_unreliable() = 0;
        const int single = cl->recognize(testImages[1]);                 // the reference's per-image call
        std::printf("\nONE%zu %d %d\nUNREL%zu %d\n", c, single, num_of_unreliable(), c, unreliable_batch);
        delete cl;
    }
    return 0;
}
