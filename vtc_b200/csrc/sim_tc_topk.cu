// sim_tc_topk.cu -- instantiates the similarity GEMM with the TopkEpi epilogue for every
// (resident query tile, cluster size) combination.  See sim_tc_kernel.cuh.
#include "sim_tc_kernel.cuh"

namespace vtc {
namespace tc {

int launch_topk(bool a_resident, int cluster, bool pair, const CUtensorMap& tmA,
                 const CUtensorMap& tmB, const Params& p, int grid, cudaStream_t s) {
  return launch_epilogue<TopkEpi>(a_resident, cluster, tmA, tmB, p, grid, s, pair);
}

}  // namespace tc
}  // namespace vtc
