// placeholder: persistent lane-refill megakernel (variant 0) -- filled in below
#pragma once
#include "uivr_kernels.cuh"
namespace uivr {
inline int launch_mega(int, bool, bool, const Params&, cudaStream_t) { return -3; }
}
