// small_m2.cu — instantiates the thread-per-environment kernels for M = 2 bodies (fp64 and fp32).
// CAV_STUB (development builds, CAVGYM_ONLY_M) leaves the table empty so that only some body counts are compiled.
#include "kernels_tma.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M2;
extern const SmallLaunchers<float> kSmallF32M2;
#ifdef CAV_STUB
const SmallLaunchers<double> kSmallF64M2 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
const SmallLaunchers<float> kSmallF32M2 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
#else
const SmallLaunchers<double> kSmallF64M2 = make_launchers<double, 2>();
const SmallLaunchers<float> kSmallF32M2 = make_launchers<float, 2>();
#endif
}  // namespace cav
