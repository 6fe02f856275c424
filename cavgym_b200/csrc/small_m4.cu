// small_m4.cu — instantiates the thread-per-environment kernels for M = 4 bodies (fp64 and fp32).
// CAV_STUB (development builds, CAVGYM_ONLY_M) leaves the table empty so that only some body counts are compiled.
#include "kernels_tma.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M4;
extern const SmallLaunchers<float> kSmallF32M4;
#ifdef CAV_STUB
const SmallLaunchers<double> kSmallF64M4 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
const SmallLaunchers<float> kSmallF32M4 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
#else
const SmallLaunchers<double> kSmallF64M4 = make_launchers<double, 4>();
const SmallLaunchers<float> kSmallF32M4 = make_launchers<float, 4>();
#endif
}  // namespace cav
