// small_m5.cu — instantiates the thread-per-environment kernels for M = 5 bodies (fp64 and fp32).
#include "kernels_small.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M5;
extern const SmallLaunchers<float> kSmallF32M5;
const SmallLaunchers<double> kSmallF64M5 = make_launchers<double, 5>();
const SmallLaunchers<float> kSmallF32M5 = make_launchers<float, 5>();
}  // namespace cav
