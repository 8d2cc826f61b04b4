// pending.cu — entry points declared in include/b2r.h whose kernels are not written yet.
// Each fails loudly with B2R_ERR_UNSUPPORTED; nothing here computes on the CPU.
#include "b2r_common.cuh"
using namespace b2r;
extern "C" {
int b2r_bm25_create(int64_t, int32_t, const int64_t*, const int32_t*, const int32_t*, const int32_t*, double,
                    const double*, double, double, int, b2r_bm25** out) {
  if (out) *out = nullptr;
  return fail(B2R_ERR_UNSUPPORTED, "b2r_bm25_create: not built yet");
}
int b2r_bm25_ndocs(const b2r_bm25*, int64_t*) { return fail(B2R_ERR_UNSUPPORTED, "bm25: not built yet"); }
int b2r_bm25_scores_host(b2r_bm25*, const int32_t*, int, double*) { return fail(B2R_ERR_UNSUPPORTED, "bm25: not built yet"); }
int b2r_bm25_topk_host(b2r_bm25*, const int32_t*, const int32_t*, int, int, double*, int64_t*) { return fail(B2R_ERR_UNSUPPORTED, "bm25: not built yet"); }
int b2r_bm25_topk_dev(b2r_bm25*, const int32_t*, const int32_t*, const int32_t*, int, int, double*, int64_t*, int64_t, void*) { return fail(B2R_ERR_UNSUPPORTED, "bm25: not built yet"); }
int b2r_bm25_destroy(b2r_bm25*) { return B2R_OK; }
int b2r_cosine_stats_dev(const float*, int64_t, int, int64_t, int64_t, double*, unsigned long long*, int, int, float*,
                         int64_t*, int, void*) { return fail(B2R_ERR_UNSUPPORTED, "cosine: not built yet"); }
}
