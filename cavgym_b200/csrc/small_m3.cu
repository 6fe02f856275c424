// small_m3.cu — instantiates the thread-per-environment kernels for M = 3 bodies (fp64 and fp32).
// CAV_STUB (development builds, CAVGYM_ONLY_M) leaves the table empty so that only some body counts are compiled.
#include "kernels_tma.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M3;
extern const SmallLaunchers<float> kSmallF32M3;
#ifdef CAV_STUB
const SmallLaunchers<double> kSmallF64M3 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
const SmallLaunchers<float> kSmallF32M3 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
#else
const SmallLaunchers<double> kSmallF64M3 = make_launchers<double, 3>();
const SmallLaunchers<float> kSmallF32M3 = make_launchers<float, 3>();
#endif
}  // namespace cav
