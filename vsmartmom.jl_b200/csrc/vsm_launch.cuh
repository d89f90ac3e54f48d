// vsm_launch.cuh -- per-(precision, tile) kernel instantiations live in their own
// translation units (vsm_inst_*.cu) so the library builds in parallel.
#pragma once
#include "vsm_column.cuh"

namespace vsm {
template <typename T, int TM>
cudaError_t launch_layer(cudaStream_t s, const ColArgs<T>& a, unsigned total, int grid, int threads, size_t smem);
template <typename T, int TM>
cudaError_t launch_adding(cudaStream_t s, const ColArgs<T>& a, unsigned total, int grid, int threads, size_t smem);
template <typename T, int TM>
void kernel_ptrs(const void** layer, const void** adding);
}  // namespace vsm
