// placeholder until the tcgen05 path lands (replaced in the next milestone)
#include "fir_common.cuh"
#include "handles.hpp"
namespace fir {
bool tensor_path_supported(int) { return false; }
int tensor_search_topk(fir_gallery*, const float*, int64_t, int, int, int32_t*, float*) { return fail(FIR_ERR_UNSUPPORTED, "tensor path not built"); }
}
