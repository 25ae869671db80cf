// placeholder until the tcgen05 path lands (replaced in a later commit)
#include "common.cuh"
using namespace qrag;
extern "C" int qrag_index_prepare(const float*, int64_t, int, uint16_t*, float*, void*) {
    return set_error(QRAG_ERR_UNSUPPORTED, "tcgen05 search path not built yet");
}
extern "C" int qrag_search_tc_workspace(int, int64_t, int, int, size_t*) {
    return set_error(QRAG_ERR_UNSUPPORTED, "tcgen05 search path not built yet");
}
extern "C" int qrag_search_topk_tc(const float*, int, const float*, const uint16_t*, const float*, int64_t, int, int,
                                   int, int64_t, double*, int64_t*, void*, size_t, void*) {
    return set_error(QRAG_ERR_UNSUPPORTED, "tcgen05 search path not built yet");
}
