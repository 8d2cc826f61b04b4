// pending.cu — entry points declared in include/b2r.h whose kernels are not written yet.
// Each fails loudly with B2R_ERR_UNSUPPORTED; nothing here computes on the CPU.
#include "b2r_common.cuh"
using namespace b2r;
extern "C" {
int b2r_cosine_stats_dev(const float*, int64_t, int, int64_t, int64_t, double*, unsigned long long*, int, int, float*,
                         int64_t*, int, void*) { return fail(B2R_ERR_UNSUPPORTED, "cosine: not built yet"); }
}
