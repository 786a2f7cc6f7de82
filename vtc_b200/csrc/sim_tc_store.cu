// sim_tc_store.cu -- instantiates the similarity GEMM with the StoreEpi epilogue (no clusters:
// these problems are small).  See sim_tc_kernel.cuh.
#include "sim_tc_kernel.cuh"

namespace vtc {
namespace tc {

int launch_store(bool a_resident, int bn, const CUtensorMap& tmA, const CUtensorMap& tmB,
                 const Params& p, int grid, cudaStream_t s) {
  if (bn == 128)  // narrow tiles: twice the CTAs for the skinny products of the CAM
    return a_resident ? launch_instance<StoreEpi, true, 1, false, 128>(tmA, tmB, p, grid, s)
                      : launch_instance<StoreEpi, false, 1, false, 128>(tmA, tmB, p, grid, s);
  return launch_epilogue_c1<StoreEpi>(a_resident, tmA, tmB, p, grid, s);
}

}  // namespace tc
}  // namespace vtc
