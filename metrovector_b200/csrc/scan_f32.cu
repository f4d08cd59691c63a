// scan_f32.cu -- instantiates the K1 scan kernels for data type f32 (one translation unit per
// data type so the build runs in parallel).
#include "scan_kernels.cuh"

namespace mvf {
namespace k {
const void* scan_kernel_f32(int op, int NQ, bool wide) { return scan_kernel_dt<0>(op, NQ, wide); }
}  // namespace k
}  // namespace mvf
