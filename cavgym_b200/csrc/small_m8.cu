// small_m8.cu — instantiates the thread-per-environment kernels for M = 8 bodies (fp64 and fp32).
// CAV_STUB (development builds, CAVGYM_ONLY_M) leaves the table empty so that only some body counts are compiled.
#include "kernels_tma.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M8;
extern const SmallLaunchers<float> kSmallF32M8;
#ifdef CAV_STUB
const SmallLaunchers<double> kSmallF64M8 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
const SmallLaunchers<float> kSmallF32M8 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
#else
const SmallLaunchers<double> kSmallF64M8 = make_launchers<double, 8>();
const SmallLaunchers<float> kSmallF32M8 = make_launchers<float, 8>();
#endif
}  // namespace cav
