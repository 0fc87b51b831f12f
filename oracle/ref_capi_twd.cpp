// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or executed from the product path.
//
// C-ABI driver around the UNMODIFIED three-way-decision classifiers of
//   /root/reference/qt_cpp/ImageTesting.cpp   (ConventionalTWDClassifier :74-186, ProposedTWDClassifier :188-288,
//                                              BruteForceClassifier :58-71)
// compiled where the file lies (see oracle/Makefile).  ImageTesting.cpp defines its own `class Classifier`, which would
// collide with classification.cpp's in the same library, so the file is textually included inside a namespace; every
// header it includes is pulled in first (their include guards turn the inner #includes into no-ops).
// Distances come from feature_distance() in the metric-specific db_features object this is linked with.
#include <vector>
#include <map>
#include <unordered_map>
#include <set>
#include <iostream>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <locale>
#include <string>
#include <chrono>
#include <functional>
#include <limits>
#include <cmath>
#include <cstdint>
#include "db.h"
#include "db_features.h"
#include <opencv2/core.hpp>
#include <opencv2/ml.hpp>

namespace fir_ref_twd {
#include "ImageTesting.cpp"
}

extern "C" {

// kind 0: ConventionalTWDClassifier(n_classes, type, th, feat_count)   type 0 Posteriors, 1 DistDiff, 2 DistRatio
// kind 1: ProposedTWDClassifier(n_classes, feat_count, th)
// kind 2: BruteForceClassifier(feat_count)
// out_class[i] = recognize(query i); out_unreliable[i] = how much num_of_unreliable grew during that call.
int fir_ref_twd_run(int kind, int type, double th, int feat_count, int n_classes, const float* g, const int32_t* labels, int64_t n, int d,
                    const float* q, int64_t nq, int32_t* out_class, uint8_t* out_unreliable) {
    using namespace fir_ref_twd;
    fir_ref_features_count = d;
    std::vector<FeaturesVector> rows(n), qrows(nq);
    std::vector<ImageInfo> db;
    db.reserve(n);
    for (int64_t j = 0; j < n; ++j) {
        rows[j].assign(g + j * d, g + (j + 1) * d);
        db.push_back(ImageInfo(labels[j], (int)j, rows[j]));
    }
    fir_ref_twd::Classifier* c = 0;
    if (kind == 0)
        c = new ConventionalTWDClassifier(n_classes, type == 0 ? ConventionalTWDClassifier::TWD_Type::Posteriors
                                                     : type == 1 ? ConventionalTWDClassifier::TWD_Type::DistDiff
                                                                 : ConventionalTWDClassifier::TWD_Type::DistRatio, th, feat_count);
    else if (kind == 1) c = new ProposedTWDClassifier(n_classes, feat_count, th);
    else if (kind == 2) c = new BruteForceClassifier(feat_count);
    else return -1;
    c->train(&db);
    for (int64_t i = 0; i < nq; ++i) {
        qrows[i].assign(q + i * d, q + (i + 1) * d);
        ImageInfo test(-1, (int)i, qrows[i]);
        const int before = num_of_unreliable;
        out_class[i] = c->recognize(test);
        if (out_unreliable) out_unreliable[i] = (uint8_t)(num_of_unreliable - before);
    }
    delete c;
    return 0;
}

}  // extern "C"
