// small_m7.cu — instantiates the thread-per-environment kernels for M = 7 bodies (fp64 and fp32).
#include "kernels_small.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M7;
extern const SmallLaunchers<float> kSmallF32M7;
const SmallLaunchers<double> kSmallF64M7 = make_launchers<double, 7>();
const SmallLaunchers<float> kSmallF32M7 = make_launchers<float, 7>();
}  // namespace cav
