// gca_launch.h - host-callable launchers of the kernels in gca_step.cu / gca_reward.cu.
#pragma once
#include <cuda_runtime.h>

#include "gca_device.cuh"

namespace gca {
cudaError_t launch_step(bool faith, bool tape, int tile, int stages, const StepArgs& a, cudaStream_t st);
cudaError_t launch_reset(bool faith, bool tape, const StepArgs& a, cudaStream_t st);
cudaError_t launch_observe(bool faith, const StepArgs& a, cudaStream_t st);
cudaError_t launch_compute_reward(const void* ag, const void* g, long long m, double radius, int kind, int is_f64,
                                  float* out, cudaStream_t st);
}  // namespace gca
