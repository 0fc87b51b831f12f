// TEST INFRASTRUCTURE ONLY.
// Force-included (g++ -include) BEFORE every reference translation unit when
// building oracle/_ref/.  The reference hard-codes its configuration in
// qt_cpp/db.h (unconditional `#define FEATURES_COUNT 1536`, db.h:86).  We claim
// that header's include guard so its body is skipped, and provide the same
// names ourselves with FEATURES_COUNT turned into a RUN-TIME value, so one
// library serves every D.  feature_distance() (db_features.cpp:22-42) receives
// start_pos/end_pos as arguments, so its code generation is unchanged.
#ifndef DB_H
#define DB_H
#define USE_CALTECH            /* db.h:11 — the reference default */
extern "C" int fir_ref_features_count;
#define FEATURES_COUNT fir_ref_features_count
extern "C" const char* fir_ref_features_file;
#define FEATURES_FILE_NAME fir_ref_features_file
#define PCA_FEATURES_FILE_NAME fir_ref_features_file
const double FRACTION = 0.03;  /* db.h:72 under USE_CALTECH */
#endif
// what QtCore drags in transitively under QT_BUILD in the reference build
#include <cmath>
#include <cfloat>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <algorithm>
#include <limits>
#include <functional>
