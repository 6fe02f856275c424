// team.cu — instantiates the team-of-warps rollout kernels (one warp per body) (kernels_team.cuh) for fp64 and fp32.
#include "kernels_team.cuh"

namespace cav {
static const TeamLaunchers<double> kTeamF64 = {launch_team_rollout<double>};
static const TeamLaunchers<float> kTeamF32 = {launch_team_rollout<float>};
template <> const TeamLaunchers<double>* team_launchers<double>() { return &kTeamF64; }
template <> const TeamLaunchers<float>* team_launchers<float>() { return &kTeamF32; }
}  // namespace cav
