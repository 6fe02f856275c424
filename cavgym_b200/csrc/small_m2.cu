// small_m2.cu — instantiates the thread-per-environment kernels for M = 2 bodies (fp64 and fp32).
#include "kernels_small.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M2;
extern const SmallLaunchers<float> kSmallF32M2;
const SmallLaunchers<double> kSmallF64M2 = make_launchers<double, 2>();
const SmallLaunchers<float> kSmallF32M2 = make_launchers<float, 2>();
}  // namespace cav
