// sim_tc.cu -- host side of the tcgen05 similarity GEMM: TMA descriptors, work planning, cluster
// choice, dispatch to the per-epilogue instantiations (sim_tc_{rank,topk,lse,store}.cu) and the
// opt-in CUDA-event timer bench.py uses for the roofline.
#include <cstdlib>
#include <mutex>
#include <utility>
#include <vector>

#include "sim_tc.cuh"

namespace vtc {
namespace tc {

int launch_rank(bool a_resident, int cluster, bool pair, const CUtensorMap& tmA,
                const CUtensorMap& tmB, const Params& p, int grid, cudaStream_t s);
int launch_topk(bool a_resident, int cluster, bool pair, const CUtensorMap& tmA,
                const CUtensorMap& tmB, const Params& p, int grid, cudaStream_t s);
int launch_lse(bool a_resident, const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p,
               int grid, cudaStream_t s);
int launch_store(bool a_resident, int bn, const CUtensorMap& tmA, const CUtensorMap& tmB,
                 const Params& p, int grid, cudaStream_t s);
int max_active_clusters_rank(int cluster);

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

int make_operand_tmap(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                      CUtensorMap* out) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return VTC_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16 || rows <= 0 || cols <= 0)
    return VTC_ERR_INVALID_ARG;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VTC_OK : VTC_ERR_DRIVER;
}

// Cluster size for the gallery multicast.  Clusters must be co-resident (persistent kernel), and a
// GPC only hosts whole clusters, so the usable SM count can shrink with the cluster size; the
// planner works with the real number from cudaOccupancyMaxActiveClusters.
static int active_clusters(int cluster) {
  static int cache[5] = {0, 0, 0, 0, 0};
  static std::mutex mu;
  if (cluster <= 1) return kNumSMs;
  std::lock_guard<std::mutex> lk(mu);
  if (cache[cluster] == 0) {
    int n = max_active_clusters_rank(cluster);
    cache[cluster] = n > 0 ? n : -1;
  }
  return cache[cluster] > 0 ? cache[cluster] : 0;
}

int choose_cluster(int64_t N, int64_t M) {
  int want = 2;  // default: pairs lose no SMs (148 = 2 x 74) and halve the L2 -> SM gallery traffic
  const char* e = getenv("VTC_CLUSTER");
  if (e && *e) want = atoi(e);
  if (want != 1 && want != 2 && want != 4) want = 1;
  const int64_t q_tiles = ceil_div<int64_t>(N, BM);
  while (want > 1 && (q_tiles < want || active_clusters(want) <= 0)) want /= 2;
  (void)M;
  return want;
}

// A cluster of two runs as a CTA pair (tcgen05 cta_group::2, one M256 MMA issued by the leader):
// each CTA stages and reads only ITS half of every gallery tile instead of receiving the full
// tile by multicast and issuing its own M128 MMAs -- a third less shared-memory operand traffic and
// half the L2 -> SM gallery bytes per SM.  Measured on B200: +4.6 % on the streamed kernel (K' > 512:
// exact mode, D = 768), +4 % at K' <= 256 (profiles/r01_summary.md); on the resident K' = 512 kernel,
// which is limited by board power, +1 % in the sustained bench loop against the multicast flavour and
// +4 % against unclustered CTAs (profiles/r02_summary.md) -- the cheaper data movement is what counts
// under the power cap.  VTC_PAIR=0/1 forces it off / on (profiling).
static bool use_pair(int num_kb) {
  static const int forced = []() {
    const char* e = getenv("VTC_PAIR");
    return e && *e ? (atoi(e) != 0 ? 1 : 0) : -1;
  }();
  (void)num_kb;
  if (forced >= 0) return forced != 0;
  return true;
}

// VTC_DBG_PROF=1 (profiling only): a small device buffer the kernel's MMA issuer and first epilogue
// warp write their wait counters to; allocated once, on first use.  The library allocates nothing
// otherwise.
constexpr int kProfWords = 256 * 8;
static unsigned long long* dbg_prof_buffer() {
  static unsigned long long* buf = []() -> unsigned long long* {
    const char* e = getenv("VTC_DBG_PROF");
    if (!e || !*e || atoi(e) == 0) return nullptr;
    unsigned long long* p = nullptr;
    if (cudaMalloc(&p, kProfWords * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, kProfWords * sizeof(unsigned long long));
    return p;
  }();
  return buf;
}
int debug_prof_read(unsigned long long* out, int max_words) {
  unsigned long long* buf = dbg_prof_buffer();
  if (!buf || !out || max_words <= 0) return VTC_ERR_INVALID_ARG;
  const int n = max_words < kProfWords ? max_words : kProfWords;
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out, buf, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemset(buf, 0, kProfWords * sizeof(unsigned long long));
  return cuda_err(e);
}

// Sub-ranges the items of the last (partial) round are cut into: one per idle cluster, at least two
// tiles each (every sub-range pays ~1 tile-time to swap the resident query tile).
static int tail_parts_for(int rem, int units, int tps, bool balance) {
  if (!balance || rem <= 0) return 1;
  int parts = units / rem;
  if (parts > tps / 2) parts = tps / 2;
  return parts < 1 ? 1 : parts;
}

Plan plan_tiles(Params& p, int max_splits, int cluster, int min_tiles_per_split, int bn,
                bool balance_tail) {
  Plan pl;
  pl.cluster = cluster < 1 ? 1 : cluster;
  pl.pair = pl.cluster == 2 && use_pair(p.num_kb);
  pl.bn = bn == 128 || bn == 64 ? bn : BN;
  static const int skip_epi = []() {
    const char* e = getenv("VTC_DBG_SKIP_EPILOGUE");
    return e && *e ? atoi(e) : 0;
  }();
  p.dbg_skip_epilogue = skip_epi;
  static const int dbg_stages = []() {
    const char* e = getenv("VTC_DBG_STAGES");
    return e && *e ? atoi(e) : 0;
  }();
  p.dbg_stages = dbg_stages;
  p.dbg_prof = dbg_prof_buffer();
  p.q_tiles = (int)ceil_div<int64_t>(p.N, BM);
  p.g_tiles = (int)ceil_div<int64_t>(p.M, pl.bn);
  const int units = pl.cluster > 1 ? active_clusters(pl.cluster) : kNumSMs;  // co-resident clusters
  const int q_groups = ceil_div(p.q_tiles, pl.cluster);
  // Items are dealt round-robin to the co-resident clusters: items / units full rounds of
  // tiles_per_split tiles (+ ~1 tile-time to swap the resident query tile) and, if items are left,
  // one more round -- of whole items, or (balance_tail) of sub-ranges that give every cluster a share.
  // Pick the split count with the lowest cost, keeping >= min_tiles_per_split tiles per item (8 for
  // the streaming epilogues, 1 for small dense products).
  const int g_tiles = p.g_tiles > 0 ? p.g_tiles : 1;
  int best = 1;
  double best_cost = 0.0;
  for (int s = 1; s <= max_splits && s <= g_tiles; ++s) {
    const int tps = ceil_div(g_tiles, s);
    if (s > 1 && tps < min_tiles_per_split) break;
    const int64_t items = (int64_t)q_groups * ceil_div(g_tiles, tps);
    const int64_t full = items / units;
    const int rem = (int)(items % units);
    double cost = (double)full * (tps + 1);
    if (rem) cost += ceil_div(tps, tail_parts_for(rem, units, tps, balance_tail)) + 1;
    if (s == 1 || cost < best_cost - 1e-9) best_cost = cost, best = s;
  }
  p.tiles_per_split = ceil_div(g_tiles, best);
  p.g_splits = ceil_div(g_tiles, p.tiles_per_split);
  const int64_t items = (int64_t)q_groups * p.g_splits;
  const int rem = (int)(items % units);
  p.tail_parts = tail_parts_for(rem, units, p.tiles_per_split, balance_tail);
  p.tail_first = p.tail_parts > 1 ? (int)(items - rem) : (int)items;
  p.num_items = p.tail_first + (p.tail_parts > 1 ? rem * p.tail_parts : 0);
  pl.grid = (int)(p.num_items < units ? p.num_items : units) * pl.cluster;
  return pl;
}

// Opt-in profiling aid (bench.py): CUDA events around every tensor-core launch on its own stream.
struct KernelTimer {
  std::mutex mu;
  bool enabled = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
};
static KernelTimer g_timer;

void kernel_timer_enable(bool on) {
  std::lock_guard<std::mutex> lk(g_timer.mu);
  g_timer.enabled = on;
}

int kernel_timer_read(double* total_ms, int* count) {
  std::lock_guard<std::mutex> lk(g_timer.mu);
  double tot = 0.0;
  int n = 0;
  for (auto& ev : g_timer.events) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(ev.second);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ev.first, ev.second);
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
    if (e != cudaSuccess) {
      g_timer.events.clear();
      return cuda_err(e);
    }
    tot += ms;
    ++n;
  }
  g_timer.events.clear();
  if (total_ms) *total_ms = tot;
  if (count) *count = n;
  return VTC_OK;
}

int launch_sim_tc(int epilogue, bool a_resident, const Plan& pl, const CUtensorMap& tmA,
                  const CUtensorMap& tmB, const Params& p, cudaStream_t s) {
  if (pl.grid <= 0) return VTC_OK;
  if (a_resident && p.num_kb > 8) return VTC_ERR_UNSUPPORTED_SHAPE;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  bool timed;
  {
    std::lock_guard<std::mutex> lk(g_timer.mu);
    timed = g_timer.enabled;
  }
  if (timed) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
  }
  int rc;
  switch (epilogue) {
    case EPI_RANK:
      rc = pl.bn == BN ? launch_rank(a_resident, pl.cluster, pl.pair, tmA, tmB, p, pl.grid, s)
                       : VTC_ERR_INVALID_ARG;
      break;
    case EPI_TOPK: rc = launch_topk(a_resident, pl.cluster, pl.pair, tmA, tmB, p, pl.grid, s); break;
    case EPI_LSE:
      rc = pl.cluster == 1 ? launch_lse(a_resident, tmA, tmB, p, pl.grid, s) : VTC_ERR_INVALID_ARG;
      break;
    case EPI_STORE:
      rc = pl.cluster == 1 ? launch_store(a_resident, pl.bn, tmA, tmB, p, pl.grid, s)
                           : VTC_ERR_INVALID_ARG;
      break;
    default: rc = VTC_ERR_INVALID_ARG;
  }
  if (timed) {
    cudaEventRecord(e1, s);
    std::lock_guard<std::mutex> lk(g_timer.mu);
    g_timer.events.emplace_back(e0, e1);
  }
  if (rc != VTC_OK) return rc;
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

}  // namespace tc
}  // namespace vtc
