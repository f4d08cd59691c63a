// scan_i8.cu -- instantiates the K1 scan kernels for data type i8 (one translation unit per
// data type so the build runs in parallel).
#include "scan_kernels.cuh"

namespace mvf {
namespace k {
const void* scan_kernel_i8(int op, int NQ, bool wide) { return scan_kernel_dt<2>(op, NQ, wide); }
}  // namespace k
}  // namespace mvf
