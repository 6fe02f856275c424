// small_m4.cu — instantiates the thread-per-environment kernels for M = 4 bodies (fp64 and fp32).
#include "kernels_small.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M4;
extern const SmallLaunchers<float> kSmallF32M4;
const SmallLaunchers<double> kSmallF64M4 = make_launchers<double, 4>();
const SmallLaunchers<float> kSmallF32M4 = make_launchers<float, 4>();
}  // namespace cav
