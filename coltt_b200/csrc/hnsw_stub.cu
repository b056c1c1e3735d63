// placeholder until hnsw.cu lands (K4); keeps every symbol of include/coltt_b200.h exported.
#include "common.cuh"
extern "C" {
COLTT_API int coltt_b200_hnsw_load(const void*, size_t, int, coltt_hnsw**) { return coltt::fail(COLTT_ERR_UNSUPPORTED, "hnsw not built yet"); }
COLTT_API void coltt_b200_hnsw_destroy(coltt_hnsw*) {}
COLTT_API int coltt_b200_hnsw_len(coltt_hnsw*, uint64_t*) { return coltt::fail(COLTT_ERR_UNSUPPORTED, "hnsw not built yet"); }
COLTT_API int coltt_b200_hnsw_search(coltt_hnsw*, const float*, size_t, int, int, uint64_t*, float*, int32_t*) { return coltt::fail(COLTT_ERR_UNSUPPORTED, "hnsw not built yet"); }
COLTT_API int coltt_b200_hnsw_last_stats(coltt_hnsw*, uint64_t*, uint64_t*) { return coltt::fail(COLTT_ERR_UNSUPPORTED, "hnsw not built yet"); }
}
