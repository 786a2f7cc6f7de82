// sim_tc_lse.cu -- instantiates the similarity GEMM with the LseEpi epilogue (no clusters:
// these problems are small).  See sim_tc_kernel.cuh.
#include "sim_tc_kernel.cuh"

namespace vtc {
namespace tc {

int launch_lse(bool a_resident, const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p,
                int grid, cudaStream_t s) {
  return launch_epilogue_c1<LseEpi>(a_resident, tmA, tmB, p, grid, s);
}

}  // namespace tc
}  // namespace vtc
