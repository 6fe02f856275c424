// small_m1.cu — instantiates the thread-per-environment kernels for M = 1 bodies (fp64 and fp32).
#include "kernels_small.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M1;
extern const SmallLaunchers<float> kSmallF32M1;
const SmallLaunchers<double> kSmallF64M1 = make_launchers<double, 1>();
const SmallLaunchers<float> kSmallF32M1 = make_launchers<float, 1>();
}  // namespace cav
