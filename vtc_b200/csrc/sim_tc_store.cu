// sim_tc_store.cu -- instantiates the similarity GEMM with the StoreEpi epilogue (no clusters:
// these problems are small).  See sim_tc_kernel.cuh.
#include "sim_tc_kernel.cuh"

namespace vtc {
namespace tc {

int launch_store(bool a_resident, int bn, const CUtensorMap& tmA, const CUtensorMap& tmB,
                 const Params& p, int grid, cudaStream_t s) {
  // always the streamed kernel: these products have a tile or two per CTA, so a resident query tile
  // buys nothing, and its shared memory is what holds the epilogue's transposing tiles
  (void)a_resident;
  // narrow tiles put more SMs to work on the skinny products of the CAM (a CTA pulls ~50 B/clk from
  // L2, so the operand bytes per CTA bound these launches, not the MMAs)
  if (bn == 128) return launch_instance<StoreEpi, false, 1, false, 128>(tmA, tmB, p, grid, s);
  if (bn == 64) return launch_instance<StoreEpi, false, 1, false, 64>(tmA, tmB, p, grid, s);
  return launch_instance<StoreEpi, false, 1>(tmA, tmB, p, grid, s);
}

}  // namespace tc
}  // namespace vtc
