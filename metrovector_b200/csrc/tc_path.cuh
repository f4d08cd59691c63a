// tc_path.cuh -- tensor-core (tcgen05) batched path, K2.  Host-side state + entry points used by
// engine.cu.  See tc_kernels.cuh for the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mvf {
namespace tc {

struct TcState {
    void* scratch = nullptr;      // candidate append buffers
    size_t scratch_bytes = 0;
    void* qprep = nullptr;        // prepared queries (operand layout) + per-query constants
    size_t qprep_bytes = 0;
    bool enabled = false;         // flipped on once the kernels are validated on hardware
};

struct SearchArgs {
    const unsigned char* rows; const float* norm; const float* sq;
    uint64_t n; uint32_t d, pitch; int dtype, metric;
    const void* queries; uint32_t nq, k; uint32_t id_add;
    unsigned long long* out_packed; uint64_t* out_ids; float* out_scores; int32_t* out_iscores;
};

inline bool eligible(const TcState& st, int path, int dtype, int qdtype, int metric, uint32_t nq, uint32_t k, uint64_t n,
                     uint32_t d, uint32_t pitch) {
    (void)path; (void)dtype; (void)qdtype; (void)metric; (void)nq; (void)k; (void)n; (void)d; (void)pitch;
    return st.enabled && false;
}
inline void search(TcState&, const SearchArgs&, cudaStream_t, int, uint64_t*) {}
inline void release(TcState& st) {
    if (st.scratch) cudaFree(st.scratch);
    if (st.qprep) cudaFree(st.qprep);
    st = TcState();
}

}  // namespace tc
}  // namespace mvf
