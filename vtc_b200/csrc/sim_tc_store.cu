// sim_tc_store.cu -- instantiates the similarity GEMM with the StoreEpi epilogue (no clusters:
// these problems are small).  See sim_tc_kernel.cuh.
#include "sim_tc_kernel.cuh"

namespace vtc {
namespace tc {

int launch_store(bool a_resident, const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p,
                int grid, cudaStream_t s) {
  return launch_epilogue_c1<StoreEpi>(a_resident, tmA, tmB, p, grid, s);
}

}  // namespace tc
}  // namespace vtc
