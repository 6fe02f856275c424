// small_m6.cu — instantiates the thread-per-environment kernels for M = 6 bodies (fp64 and fp32).
#include "kernels_small.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M6;
extern const SmallLaunchers<float> kSmallF32M6;
const SmallLaunchers<double> kSmallF64M6 = make_launchers<double, 6>();
const SmallLaunchers<float> kSmallF32M6 = make_launchers<float, 6>();
}  // namespace cav
