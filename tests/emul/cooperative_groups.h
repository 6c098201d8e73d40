/* tests/emul/cooperative_groups.h — TEST INFRASTRUCTURE: the grid group of a cooperative launch on the lock-step model
 * (cuda_runtime.h, emu.cpp): sync() is a rendezvous of every live thread of every block. */
#pragma once
#include "cuda_runtime.h"
namespace cooperative_groups {
struct grid_group { void sync() const { emu::grid_sync(); } };
inline grid_group this_grid(){ return grid_group(); }
}
