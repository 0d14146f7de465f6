// pus_span.cu -- the second instantiation of the device code (pus::kspan): the persistent kernel with the hooks for one
// graph spanning several ranks (mirrored stores into the peers' arenas, cross-rank barrier).  Kept
// in its own translation unit so that it compiles in parallel with pus_engine.cu (which holds pus::kplain, the
// single-GPU kernel with those hooks compiled out).  The host engine reaches it through the two functions below.
#include <cuda_runtime.h>

#include "pus_graph.hpp"

namespace pus {
namespace kspan {
#include "pus_kernels.cuh"
#include "pus_driver.cuh"
}  // namespace kspan

void* span_kernel_ptr() { return (void*)kspan::lm_kernel; }
size_t span_devgraph_bytes() { return sizeof(kspan::DevGraph); }
cudaError_t span_kernel_prepare() {
  return cudaFuncSetAttribute(kspan::lm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kspan::kSmemBytes);
}
}  // namespace pus
