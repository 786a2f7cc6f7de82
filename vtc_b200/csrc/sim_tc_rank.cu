// sim_tc_rank.cu -- instantiates the similarity GEMM with the RankEpi epilogue for every
// (resident query tile, cluster size) combination.  See sim_tc_kernel.cuh.
#include "sim_tc_kernel.cuh"

namespace vtc {
namespace tc {

int launch_rank(bool a_resident, int cluster, bool pair, const CUtensorMap& tmA,
                 const CUtensorMap& tmB, const Params& p, int grid, cudaStream_t s) {
  return launch_epilogue<RankEpi>(a_resident, cluster, tmA, tmB, p, grid, s, pair);
}

int max_active_clusters_rank(int cluster) {
  if (cluster == 4) return max_active_clusters<RankEpi, true, 4>();
  if (cluster == 2) return max_active_clusters<RankEpi, true, 2>();
  return kNumSMs;
}

}  // namespace tc
}  // namespace vtc
