// small_m8.cu — instantiates the thread-per-environment kernels for M = 8 bodies (fp64 and fp32).
#include "kernels_small.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M8;
extern const SmallLaunchers<float> kSmallF32M8;
const SmallLaunchers<double> kSmallF64M8 = make_launchers<double, 8>();
const SmallLaunchers<float> kSmallF32M8 = make_launchers<float, 8>();
}  // namespace cav
