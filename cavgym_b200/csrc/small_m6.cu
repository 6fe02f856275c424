// small_m6.cu — instantiates the thread-per-environment kernels for M = 6 bodies (fp64 and fp32).
// CAV_STUB (development builds, CAVGYM_ONLY_M) leaves the table empty so that only some body counts are compiled.
#include "kernels_tma.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M6;
extern const SmallLaunchers<float> kSmallF32M6;
#ifdef CAV_STUB
const SmallLaunchers<double> kSmallF64M6 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
const SmallLaunchers<float> kSmallF32M6 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
#else
const SmallLaunchers<double> kSmallF64M6 = make_launchers<double, 6>();
const SmallLaunchers<float> kSmallF32M6 = make_launchers<float, 6>();
#endif
}  // namespace cav
