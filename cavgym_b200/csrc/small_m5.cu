// small_m5.cu — instantiates the thread-per-environment kernels for M = 5 bodies (fp64 and fp32).
// CAV_STUB (development builds, CAVGYM_ONLY_M) leaves the table empty so that only some body counts are compiled.
#include "kernels_tma.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M5;
extern const SmallLaunchers<float> kSmallF32M5;
#ifdef CAV_STUB
const SmallLaunchers<double> kSmallF64M5 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
const SmallLaunchers<float> kSmallF32M5 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
#else
const SmallLaunchers<double> kSmallF64M5 = make_launchers<double, 5>();
const SmallLaunchers<float> kSmallF32M5 = make_launchers<float, 5>();
#endif
}  // namespace cav
