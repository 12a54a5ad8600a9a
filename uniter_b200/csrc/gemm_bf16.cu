// bf16 instantiations of the tcgen05 GEMM core (split from fp16 for build parallelism).
#include "gemm_impl.cuh"
namespace ub {
int gemm_dispatch_bf16(int bn, int cluster, int a_major, int b_major, const GemmParams& p,
                       const CUtensorMap& tmA, const CUtensorMap& tmB, int grid, cudaStream_t stream) {
  return gemm_dispatch<true>(bn, cluster, a_major, b_major, p, tmA, tmB, grid, stream);
}
int gemm_group_dispatch_bf16(const void* tm, const GroupedParams& g, int grid, cudaStream_t stream) {
  return gemm_group_dispatch<true>(*reinterpret_cast<const TmPack*>(tm), g, grid, stream);
}
}  // namespace ub
