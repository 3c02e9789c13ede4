// deepz_launch.h -- launchers of the deepz_kernel instantiations, one translation unit per KSREG.
#pragma once
#include "deepz_types.cuh"

#define DZ_LAUNCH_ARGS                                                                                          \
    bool global, bool stats, bool full, bool wide, int grid, int threads, size_t smem, cudaStream_t stream, const DevNet *net,        \
        const DeepzInput &in, const DeepzOutput &out, DevCallState *st, const DeepzSched &sp
cudaError_t launch_deepz_ks8(DZ_LAUNCH_ARGS);
cudaError_t launch_deepz_ks16(DZ_LAUNCH_ARGS);
cudaError_t launch_deepz_ks25(DZ_LAUNCH_ARGS);
cudaError_t launch_deepz_ks32(DZ_LAUNCH_ARGS);
