// TEST INFRASTRUCTURE: parse-only stand-in for <cooperative_groups.h> (see cuda_runtime.h in this directory).
#pragma once
namespace cooperative_groups {
struct grid_group { void sync() const {} };
inline grid_group this_grid() { return grid_group{}; }
}  // namespace cooperative_groups
