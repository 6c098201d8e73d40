/* tests/emul/cooperative_groups.h — TEST INFRASTRUCTURE: ns_par_dev.cuh includes <cooperative_groups.h> for the grid-scope engine's
 * grid_group; the lock-step model only compiles the CTA scope (NS_SCOPE_CTA), where the type is never used. */
#pragma once
namespace cooperative_groups { struct grid_group { void sync(); }; }
