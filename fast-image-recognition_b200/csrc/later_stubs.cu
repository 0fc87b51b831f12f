// placeholders for the fp64 classifier and DEM entry points (replaced by cls_kernels.cu / dem_kernels.cu)
#include "fir_common.cuh"
using namespace fir;
extern "C" {
int fir_classifier_create(const double*, const int32_t*, int64_t, int32_t, int32_t, const double*, fir_classifier**) { return fail(FIR_ERR_UNSUPPORTED, "not built yet"); }
int fir_classifier_destroy(fir_classifier*) { return FIR_OK; }
int fir_classifier_knn(fir_classifier*, const double*, int64_t, int32_t, int32_t*) { return fail(FIR_ERR_UNSUPPORTED, "not built yet"); }
int fir_classifier_pnn(fir_classifier*, const double*, int64_t, double*, int32_t*) { return fail(FIR_ERR_UNSUPPORTED, "not built yet"); }
int fir_dem_build(fir_gallery*, const fir_dem_params*, fir_dem**) { return fail(FIR_ERR_UNSUPPORTED, "not built yet"); }
int fir_dem_destroy(fir_dem*) { return FIR_OK; }
int fir_dem_info(const fir_dem*, int32_t*, int32_t*, float*) { return fail(FIR_ERR_UNSUPPORTED, "not built yet"); }
int fir_dem_get_pivots(const fir_dem*, int32_t*) { return fail(FIR_ERR_UNSUPPORTED, "not built yet"); }
int fir_dem_get_pivot_matrix(const fir_dem*, float*) { return fail(FIR_ERR_UNSUPPORTED, "not built yet"); }
int fir_dem_get_min_other(const fir_dem*, float*) { return fail(FIR_ERR_UNSUPPORTED, "not built yet"); }
int fir_dem_search(fir_dem*, const float*, int64_t, int32_t, int32_t, int32_t*, float*, uint8_t*, int32_t*) { return fail(FIR_ERR_UNSUPPORTED, "not built yet"); }
}
