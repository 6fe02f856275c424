// small_m3.cu — instantiates the thread-per-environment kernels for M = 3 bodies (fp64 and fp32).
#include "kernels_small.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M3;
extern const SmallLaunchers<float> kSmallF32M3;
const SmallLaunchers<double> kSmallF64M3 = make_launchers<double, 3>();
const SmallLaunchers<float> kSmallF32M3 = make_launchers<float, 3>();
}  // namespace cav
