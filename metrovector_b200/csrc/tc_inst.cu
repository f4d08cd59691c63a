// tc_inst.cu -- instantiates the K2 kernels of tc_kernels.cuh for ONE operand kind and ONE epilogue-group count
// (-DTC_KIND=0|1|2 -DTC_EG=2|4; see the Makefile), so the five objects compile in parallel.
#include "tc_kernels.cuh"

#ifndef TC_KIND
#error "compile with -DTC_KIND=0|1|2 -DTC_EG=2|4"
#endif

namespace mvf {
namespace tc {

template <int KIND, int MODE, int EG>
static const void* pick_form(int cl, bool pair) {
    if (pair) return (const void*)gemm_topk_tc<KIND, MODE, 2, EG, true>;
    if constexpr (KIND == KIND_TF32) {
        return (const void*)gemm_topk_tc<KIND, MODE, 1, EG, false>;      // split stages: pair form or single CTAs only
    } else {
        return cl == 2 ? (const void*)gemm_topk_tc<KIND, MODE, 2, EG, false> : (const void*)gemm_topk_tc<KIND, MODE, 1, EG, false>;
    }
}

#define TC_CAT2(a, b, c) a##b##_eg##c
#define TC_CAT(a, b, c) TC_CAT2(a, b, c)
// kernel_k<KIND>_eg<EG>(mode, cl, pair): MODE 0 ranks |x|^2 - 2 q.x (L2), MODE 1 ranks -q.x / |x| or -q.x, MODE 3 is the
// integer kinds' cosine (approximate common path, exact keys for the candidates)
const void* TC_CAT(kernel_k, TC_KIND, TC_EG)(int mode, int cl, bool pair) {
#if TC_KIND == 1
    if (mode == 3) return pick_form<TC_KIND, 3, TC_EG>(cl, pair);
#endif
    return mode == 0 ? pick_form<TC_KIND, 0, TC_EG>(cl, pair) : pick_form<TC_KIND, 1, TC_EG>(cl, pair);
}

}  // namespace tc
}  // namespace mvf
