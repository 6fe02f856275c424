// small_m7.cu — instantiates the thread-per-environment kernels for M = 7 bodies (fp64 and fp32).
// CAV_STUB (development builds, CAVGYM_ONLY_M) leaves the table empty so that only some body counts are compiled.
#include "kernels_tma.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M7;
extern const SmallLaunchers<float> kSmallF32M7;
#ifdef CAV_STUB
const SmallLaunchers<double> kSmallF64M7 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
const SmallLaunchers<float> kSmallF32M7 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
#else
const SmallLaunchers<double> kSmallF64M7 = make_launchers<double, 7>();
const SmallLaunchers<float> kSmallF32M7 = make_launchers<float, 7>();
#endif
}  // namespace cav
