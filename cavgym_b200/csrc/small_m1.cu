// small_m1.cu — instantiates the thread-per-environment kernels for M = 1 bodies (fp64 and fp32).
// CAV_STUB (development builds, CAVGYM_ONLY_M) leaves the table empty so that only some body counts are compiled.
#include "kernels_tma.cuh"

namespace cav {
extern const SmallLaunchers<double> kSmallF64M1;
extern const SmallLaunchers<float> kSmallF32M1;
#ifdef CAV_STUB
const SmallLaunchers<double> kSmallF64M1 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
const SmallLaunchers<float> kSmallF32M1 = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
#else
const SmallLaunchers<double> kSmallF64M1 = make_launchers<double, 1>();
const SmallLaunchers<float> kSmallF32M1 = make_launchers<float, 1>();
#endif
}  // namespace cav
