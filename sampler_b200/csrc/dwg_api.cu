// dwg_api.cu -- the C ABI (include/dwg.h) over the sweep kernels.
//
// dwg_graph  = a GraphImage (host) ; dwg_handle = one replica on one GPU.
// All device work of a handle goes to its own non-blocking stream; sweeps are
// asynchronous, getters synchronise.  No CPU fallback exists: every compute entry point
// needs a CUDA device and reports DWG_ERR_CUDA without one.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/dwg.h"
#include "dwg_kernels.cuh"
#include "dwg_sell.cuh"
#include "dwg_cell.cuh"
#include "graph_image.h"
#include "host/parallel.h"

using namespace dwg;

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define CU(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail(DWG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
  } while (0)

struct dwg_graph {
  GraphImage img;
  // the big streams of the image are page-locked on the first dwg_create so that this and every
  // later upload of the graph (one per replica / per run) is a DMA at the PCIe rate instead of
  // a bounce through the driver's staging buffers; undone by dwg_graph_free
  mutable std::mutex pin_mu;
  mutable std::vector<const void *> pinned;
  mutable bool pin_wanted = true;  // dwg_graph_set_page_locking
  void pin() const {
    std::lock_guard<std::mutex> lk(pin_mu);
    if (!pinned.empty() || !pin_wanted) return;
    auto reg = [&](const void *p, size_t bytes) {
      if (bytes < (1u << 20)) return;
      if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) pinned.push_back(p);
      else cudaGetLastError();  // not fatal: the copy falls back to pageable memory
    };
    if (img.sell || img.cell) reg(img.blob.data(), img.blob.size() * 4);
    if (!img.sell) {
      reg(img.e_wid.data(), img.e_wid.size() * 4);
      reg(img.e_tag.data(), img.e_tag.size());
      reg(img.oth.data(), img.oth.size() * 4);
      reg(img.row_ptr.data(), img.row_ptr.size() * 4);
      reg(img.e_fv32.data(), img.e_fv32.size() * 4);
      reg(img.e_fv64.data(), img.e_fv64.size() * 8);
    }
    reg(img.row_of_value.data(), img.row_of_value.size() * 4);
    reg(img.var_flags.data(), img.var_flags.size());
    reg(img.evid_val8.data(), img.evid_val8.size());
    reg(img.sl_off.data(), img.sl_off.size() * 4);
    if (pinned.empty()) pinned.push_back(nullptr);  // tried
  }
  // per colour class: the first slice from which every slice holds evidence / padding only
  // (slices are sorted non-evidence first inside a class, compile.cc); computed once per graph
  mutable std::vector<uint32_t> cls_query_end, cls_max_words;  // the latter: longest slice of the class
  const std::vector<uint32_t> &query_end() const {
    std::lock_guard<std::mutex> lk(pin_mu);
    if (!cls_query_end.empty() || !(img.sell || img.cell)) return cls_query_end;
    const uint32_t nc = img.ncolor();
    cls_query_end.assign(nc, 0);
    cls_max_words.assign(nc, 0);
    for (uint32_t c = 0; c < nc; ++c) {
      const uint32_t sb = img.color_start[c] / 32, se = img.color_start[c + 1] / 32;
      for (uint32_t sl = sb; sl < se; ++sl)
        cls_max_words[c] = std::max(cls_max_words[c], (img.sl_off[sl + 1] - img.sl_off[sl]) * 8u);
      uint32_t qe = sb;
      for (uint32_t sl = se; sl > sb && qe == sb; --sl) {  // from the evidence tail backwards
        bool query = false;
        for (uint32_t d = (sl - 1) * 32; d < sl * 32 && !query; ++d)
          query = !(img.var_flags[d] & (VF_EVID | VF_PAD));
        if (query) qe = sl;
      }
      cls_query_end[c] = qe;
    }
    return cls_query_end;
  }
  ~dwg_graph() {
    for (const void *p : pinned)
      if (p) cudaHostUnregister(const_cast<void *>(p));
  }
};


namespace {
// dwg_create records every graph array it puts on the device (pointer, bytes) and, for a replica
// created with defer_upload, skips the copies: a peer replica's arrays arrive later (dwg_graph_segments)
struct UploadScope {
  std::vector<std::pair<void *, size_t>> *segments = nullptr;
  bool defer = false;
};
thread_local UploadScope g_upload;

// Device memory comes from the device's default stream-ordered pool with an unlimited release
// threshold: what a destroyed handle frees stays mapped in the process, so the next dwg_create (a
// replica per inference request, the repetitions of a benchmark) does not pay the driver for mapping
// 0.6 GB again -- that cost was 20 to 130 ms at random, more than the 12 ms upload it precedes.
// Allocation and release are ordered on the legacy default stream and completed before returning
// (the handles work on their own non-blocking streams); DWG_NO_POOL=1 restores cudaMalloc.
static bool dev_pool_on(int dev) {
  static std::mutex mu;
  static std::vector<int> state;  // per device: 0 unknown, 1 pool, 2 plain
  std::lock_guard<std::mutex> lock(mu);
  if ((int)state.size() <= dev) state.resize(dev + 1, 0);
  if (!state[dev]) {
    int ok = 0;
    cudaDeviceGetAttribute(&ok, cudaDevAttrMemoryPoolsSupported, dev);
    const char *e = getenv("DWG_NO_POOL");
    cudaMemPool_t pool = nullptr;
    if (ok && !(e && *e == '1') && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = ~0ull;
      ok = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess;
    } else ok = 0;
    cudaGetLastError();
    state[dev] = ok ? 1 : 2;
  }
  return state[dev] == 1;
}
static cudaError_t dev_alloc(void **p, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (!dev_pool_on(dev)) return cudaMalloc(p, bytes);
  cudaError_t e = cudaMallocAsync(p, bytes, (cudaStream_t)0);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize((cudaStream_t)0);
}
static void dev_free(void *p) {
  cudaPointerAttributes a{};
  int dev = 0;
  if (cudaPointerGetAttributes(&a, p) == cudaSuccess) dev = a.device;
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != dev) cudaSetDevice(dev);
  if (dev_pool_on(dev)) {
    cudaFreeAsync(p, (cudaStream_t)0);
    cudaStreamSynchronize((cudaStream_t)0);
  } else cudaFree(p);
  if (cur != dev) cudaSetDevice(cur);
}

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  ~DevBuf() { reset(); }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  void reset() {
    if (p) dev_free(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    n = count;
    // 16 B of slack so tile-granular (16 B) bulk copies may overrun the logical end
    return dev_alloc((void **)&p, count * sizeof(T) + 16);
  }
  template <typename S>
  cudaError_t upload(const std::vector<S> &v, cudaStream_t s) {
    static_assert(sizeof(S) == sizeof(T), "");
    cudaError_t e = alloc(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    if (g_upload.segments) g_upload.segments->emplace_back((void *)p, v.size() * sizeof(T));
    if (g_upload.defer) return cudaSuccess;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
  size_t bytes() const { return n * sizeof(T); }
};
}  // namespace

namespace {
// Page-locked landing buffer of the getters, one per process: page-locking costs more than the
// transfers it serves, so it is paid once, not per handle.  Held for the duration of a download.
struct LandingBuffer {
  std::mutex mu;
  uint32_t *buf = nullptr;
  size_t words = 0;
  uint32_t *get(size_t w) {
    if (w > words) {
      if (buf) cudaFreeHost(buf);
      buf = nullptr; words = 0;
      if (cudaHostAlloc((void **)&buf, w * 4, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      words = w;
    }
    return buf;
  }
};
LandingBuffer g_landing;
}  // namespace

struct dwg_handle {
  const GraphImage *img = nullptr;
  dwg_opts_t opts{};
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  // per-variable arrays in schedule order (nvar_pad)
  DevBuf<uint8_t> var_flags, evid_val8;
  DevBuf<uint32_t> evid_val32, var_row;
  // CSR stream: always for categorical graphs; for sliced-ELL graphs only on demand
  // (parity hooks, sequential debug mode)
  bool csr_ready = false;
  DevBuf<uint8_t> e_tag;
  DevBuf<uint32_t> row_ptr, row_oth, e_wid, oth;
  DevBuf<float> e_fv32;
  DevBuf<double> e_fv64, truthiness;
  // sliced-ELL stream (Boolean graphs, parallel schedules)
  DevBuf<uint32_t> sl_off, blob;
  // weights (internal order) and the batch-update statics
  DevBuf<uint8_t> w_fixed;
  DevBuf<double> weights, prev_weights, grad_acc, norms;
  DevBuf<float> cls_N, cls_Q, cls_H;
  bool learns = true;  // false: no variable takes an SGD step under these options
  // state
  DevBuf<uint8_t> assign_free8, assign_evid8;
  DevBuf<uint32_t> assign_free32, assign_evid32, tallies;
  DevBuf<double> r_free, r_evid;  // parity hook
  bool explicit_r = false, sequential = false;
  // the last operation enqueued on the stream was a staged inference launch of this handle (the
  // next one may be a programmatic dependent launch); cleared by every other entry point
  bool chain_open = false;
  uint64_t sweep_counter = 0, infer_sweeps = 0;
  std::vector<uint64_t> extra_nsamples;  // from dwg_add_tallies
  DevGraph dg{};
  DevState ds{};
  SellGraph sg{};
  SellState ss{};
  int sell_blocks_infer = 0, sell_blocks_learn = 0;
  // bulk-async staged inference kernel (dwg_sell.cuh): ring depth, stage capacity in words,
  // grid cap; 0 stages = not available (DWG_SELL_DIRECT=1, or no room for two stages)
  uint32_t sell_stages = 0, sell_stage_words = 0, sell_staged_warps = 0;
  int sell_blocks_staged = 0;
  // categorical sliced-ELL inference kernel (dwg_cell.cuh)
  CellGraph cg{};
  uint32_t cell_stages = 0, cell_stage_words = 0;
  int cell_blocks = 0;
  // categorical learning kernel (k_cell_learn_live): whole-graph update counts / curvature bounds
  // per weight, hot weights kept in shared memory (a prefix of the internal order)
  bool cell_learn = false;
  DevBuf<float> lv_N, lv_H;
  // per colour class: the first slice from which every slice holds evidence / padding only
  // (slices are sorted non-evidence first inside a class, compile.cc)
  std::vector<uint32_t> cls_query_end;
  // learning sweeps of sliced-ELL graphs (k_sell_learn_live): visit order, per-weight statistics
  // in internal weight order (device: cls_N / cls_H; host: curvatures sorted for the epoch plan)
  DevBuf<uint32_t> learn_visit;
  DevBuf<double> part_sums;        // hubs cut into parts (SF_PART): per-slice sums and counters
  DevBuf<uint32_t> part_done;
  std::vector<float> learn_H_sorted;   // curvature bounds of the learnable weights, ascending
  double learn_rho = 1.0;              // coupling: learnable records per learning variable / 10
  uint32_t learn_stage_words = 8;      // capacity of one ring stage of the learning kernel
  size_t device_bytes = 0;
  std::vector<std::pair<void *, size_t>> segments;  // the graph arrays on the device (dwg_graph_segments)
  // tally download: device row of every (variable, value) in ABI order (uploaded with the graph) and
  // the gathered tallies; the page-locked landing pieces are the process's (g_landing)
  DevBuf<uint32_t> row_of_value, tallies_abi, tallies_packed;
  DevBuf<unsigned long long> nsamples_dev;  // dwg_get_tallies*: the sample counts for a page-locked destination
  DevBuf<uint8_t> evid_abi;                 // Variable::is_evid in ABI order (only when evidence is not sampled)
  cudaEvent_t piece_done[2] = {nullptr, nullptr};
  // row_of_value is read by the tally getters only: its host -> device copy is queued on `late`
  // once the graph's other copies have landed (so that it does not share the link with them) and
  // rides under the first sweeps; the getters wait for late_done
  cudaStream_t late = nullptr;
  cudaEvent_t late_done = nullptr;
  bool row_map_ready = false;
  ~dwg_handle() {
    if (stream) cudaStreamSynchronize(stream);  // (a failed dwg_create may leave copies in flight)
    for (cudaEvent_t e : piece_done)
      if (e) cudaEventDestroy(e);
    if (late_done) cudaEventDestroy(late_done);
    if (late) { cudaStreamSynchronize(late); cudaStreamDestroy(late); }
    if (stream) cudaStreamDestroy(stream);
  }
};

// ---- sliced-ELL kernel dispatch over (multi-variable factors?, feature-value mode) ----
namespace {
template <typename K>
int sell_grid(K kernel, size_t smem, int num_sms, int *blocks, int threads = SELL_THREADS) {
  // (the 48 KB default covers dynamic + static shared memory: a kernel with a few static words and
  // exactly 48 KB of dynamic memory needs the opt-in too -- C5's learning ring landed there)
  if (smem > 40 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  if (per_sm < 1) {
    cudaFuncAttributes fa{};
    cudaFuncGetAttributes(&fa, kernel);
    return fail(DWG_ERR_CUDA, "sweep kernel does not fit on an SM (" + std::to_string(threads) + " threads x " + std::to_string(fa.numRegs) +
                                  " registers, " + std::to_string(smem) + " + " + std::to_string(fa.sharedSizeBytes) + " B of shared memory)");
  }
  *blocks = per_sm * num_sms;
  return DWG_OK;
}
template <bool MULTI, int FV, bool WIDE>
int sell_setup_t(dwg_handle *h) {
  const size_t n_hot = h->img->n_hot;
  int rc = sell_grid(k_sell_infer<MULTI, FV, WIDE>, 8 * n_hot, h->num_sms, &h->sell_blocks_infer);
  if (rc) return rc;
  {  // learning: a two-stage ring per warp of whatever fits beside the weight tables (longer slices are read from HBM)
    const size_t fixed = sell_learn_ring_off((uint32_t)n_hot), lbudget = 112 * 1024;
    uint32_t cap = fixed + 1024 < lbudget ? (uint32_t)((lbudget - fixed) / ((SELL_LEARN_THREADS / 32) * SELL_LEARN_STAGES * 4)) & ~3u : 8u;
    if (const char *e = getenv("DWG_SELL_STAGE_WORDS")) cap = std::min<uint32_t>(cap, std::max(8, atoi(e)) & ~3u);  // tests
    h->learn_stage_words = std::max(8u, std::min(cap, (h->img->sell_max_words + 3u) & ~3u));
    rc = sell_grid(k_sell_learn_live<MULTI, FV, WIDE>, sell_learn_smem((uint32_t)n_hot, h->learn_stage_words), h->num_sms,
                   &h->sell_blocks_learn, SELL_LEARN_THREADS);
    if (rc) return rc;
  }
  // staged (bulk-async) inference kernel: ring geometry for this graph
  h->sell_stages = 0;
  const char *direct = getenv("DWG_SELL_DIRECT");
  if (direct && *direct == '1') return DWG_OK;
  size_t budget = 112 * 1024;  // per CTA, two per SM out of 228 KB (1 KB reserved each)
  // measured on C2 / C5 (profiles/r1_v4_ab.txt): with neighbour gathers in the loop two deep stages
  // beat three or four shallow ones (a deeper prefetch queues the latency-critical L2 gathers behind
  // bulk traffic), unary-only graphs gain from the deeper ring
  uint32_t max_stages = MULTI ? 2 : SELL_MAX_STAGES;
  if (const char *e = getenv("DWG_SELL_STAGES")) max_stages = std::max(2, std::min(SELL_MAX_STAGES, atoi(e)));
  // measured (profiles/r1_v4_ab.txt): with neighbour gathers in the loop 8 warps x <= 128 registers per
  // CTA win (more loads in flight per warp beat more warps), unary-only graphs stream best with 16 x 64
  uint32_t nwarp = MULTI ? 8 : 16;
  if (const char *e = getenv("DWG_SELL_WARPS")) nwarp = atoi(e) == 16 ? 16 : atoi(e) == 12 ? 12 : atoi(e) == 4 ? 4 : 8;
  const uint32_t n_tab = (uint32_t)std::min<uint64_t>({h->img->nweight, n_hot, SELL_MAX_TAB});
  // the deepest ring (<= SELL_MAX_STAGES, two CTAs per SM) whose stage holds at least 99.5 % of the
  // slices whole (90 % for the two-stage ring); the rare longer slice is staged up to its stage's
  // end and read from HBM directly.  Long slices (20 f64-valued features = 7.7 KB) get fewer warps
  // per CTA (8, then 4) so that two stages still fit.
  const std::vector<uint32_t> &off = h->img->sl_off;
  const size_t nslice = off.empty() ? 0 : off.size() - 1;
  const uint32_t want = (h->img->sell_max_words + 3u) & ~3u;
  uint32_t stages = 0, words = 0;
  for (; !stages; nwarp = nwarp == 12 ? 8 : nwarp / 2) {
    for (uint32_t st = max_stages; st >= 2 && !stages; --st) {
      const size_t fixed = sell_staged_ring_off(n_tab, st, nwarp);
      uint32_t cap = (uint32_t)((budget - fixed) / (nwarp * st * 4)) & ~3u;
      if (const char *e = getenv("DWG_SELL_STAGE_WORDS")) cap = std::min<uint32_t>(cap, std::max(8, atoi(e)) & ~3u);  // tests
      size_t fit = 0;
      // (slices of a streamed synthetic image all have the same length: look at a sample)
      const size_t stride = nslice > (1u << 22) ? nslice >> 20 : 1;
      size_t seen = 0;
      for (size_t i = 0; i < nslice; i += stride, ++seen) fit += (off[i + 1] - off[i]) * 8u <= cap;
      if (fit * 1000 >= seen * 995 || (st == 2 && fit * 10 >= seen * 9)) { stages = st; words = std::min(cap, want); }
    }
    // measured (profiles/r2_c4_ab.txt, 20 f64 features = 7.7 KB per slice, 20 M variables): four
    // warps per CTA with a two-stage ring sweep in 2.88 ms, the direct kernel (32 warps per SM, loads
    // issued a batch ahead) in 1.08 ms -- eight resident warps cannot cover the latency of the
    // weight gathers.  So the ladder stops at eight warps unless DWG_SELL_WARPS=4 asks for the rest.
    if (stages || nwarp == 4 || (nwarp == 8 && !getenv("DWG_SELL_WARPS"))) break;
  }
  // slices too long for the ring: the direct kernel streams them from HBM
  if (!stages) return DWG_OK;
  const size_t smem = sell_staged_smem(n_tab, stages, words, nwarp);
  const bool allhot = h->img->nweight <= n_tab;
#define STAGED_GRID(AH, NT) sell_grid(k_sell_infer_staged<MULTI, FV, AH, NT, WIDE>, smem, h->num_sms, &h->sell_blocks_staged, NT)
  rc = nwarp == 16 ? (allhot ? STAGED_GRID(true, 512) : STAGED_GRID(false, 512))
       : nwarp == 12 ? (allhot ? STAGED_GRID(true, 384) : STAGED_GRID(false, 384))
       : nwarp == 8 ? (allhot ? STAGED_GRID(true, 256) : STAGED_GRID(false, 256))
                    : (allhot ? STAGED_GRID(true, 128) : STAGED_GRID(false, 128));
#undef STAGED_GRID
  if (rc) return rc;
  h->sell_staged_warps = nwarp;
  h->sg.n_tab = n_tab;
  h->sell_stages = stages;
  h->sell_stage_words = words;
  return DWG_OK;
}
template <bool MULTI, int FV, bool WIDE>
// staged: the bulk-async inference kernel; pdl: programmatic dependent launch behind the previous
// inference launch
void sell_launch_t(dwg_handle *h, bool staged, bool pdl, uint32_t blocks, const SellGraph &sg, const SweepParams &p) {
  const size_t n_hot = h->img->n_hot;
  if (staged) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(h->sell_staged_warps * 32);
    cfg.dynamicSmemBytes = sell_staged_smem(sg.n_tab, h->sell_stages, h->sell_stage_words, h->sell_staged_warps);
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;  // only behind another launch of this kernel (see launch_sweep)
    const bool allhot = h->img->nweight <= sg.n_tab;
#define STAGED_LAUNCH(AH, NT) \
  cudaLaunchKernelEx(&cfg, k_sell_infer_staged<MULTI, FV, AH, NT, WIDE>, sg, h->ss, p, h->sell_stages, h->sell_stage_words)
    if (h->sell_staged_warps == 16) { if (allhot) STAGED_LAUNCH(true, 512); else STAGED_LAUNCH(false, 512); }
    else if (h->sell_staged_warps == 12) { if (allhot) STAGED_LAUNCH(true, 384); else STAGED_LAUNCH(false, 384); }
    else if (h->sell_staged_warps == 8) { if (allhot) STAGED_LAUNCH(true, 256); else STAGED_LAUNCH(false, 256); }
    else { if (allhot) STAGED_LAUNCH(true, 128); else STAGED_LAUNCH(false, 128); }
#undef STAGED_LAUNCH
  } else k_sell_infer<MULTI, FV, WIDE><<<blocks, SELL_THREADS, 8 * n_hot, h->stream>>>(sg, h->ss, p);
}
// live: no shared-memory copy of the weights at all (the narrow tail of a learning sweep: every
// record reads its weight from L2 and pushes its step there, nothing to refresh or flush)
template <bool MULTI, int FV, bool WIDE>
void sell_learn_launch_t(dwg_handle *h, uint32_t blocks, uint32_t threads, bool live, const SweepParams &p,
                         const SellLearn &L) {
  SellGraph sg = h->sg;
  if (live) sg.n_hot = 0;
  k_sell_learn_live<MULTI, FV, WIDE><<<blocks, threads, sell_learn_smem(sg.n_hot, h->learn_stage_words), h->stream>>>(
      sg, h->ss, p, L, h->learn_stage_words);
}
#ifdef DWG_DEV_BUILD  // development builds (DWG_DEV_BUILD=1 python -m sampler_b200._build): only the C2-shaped kernels, a third of the compile time
#define SELL_SWITCH(CALL) { (void)m; CALL(true, FV_ONE, false); }
#else
#define SELL_SWITCH_W(CALL, WD)                                                \
  if (m.has_multi) {                                                           \
    if (m.fv_mode == FV_ONE) { CALL(true, FV_ONE, WD); }                       \
    else if (m.fv_mode == FV_F32) { CALL(true, FV_F32, WD); }                  \
    else { CALL(true, FV_F64, WD); }                                           \
  } else {                                                                     \
    if (m.fv_mode == FV_ONE) { CALL(false, FV_ONE, WD); }                      \
    else if (m.fv_mode == FV_F32) { CALL(false, FV_F32, WD); }                 \
    else { CALL(false, FV_F64, WD); }                                          \
  }
// (graphs with warp-per-variable slices run kernels instantiated with the code for them)
#define SELL_SWITCH(CALL)                                                      \
  if (m.has_wide) { SELL_SWITCH_W(CALL, true) } else { SELL_SWITCH_W(CALL, false) }
#endif
constexpr int CELL_THREADS = 256;
int cell_setup(dwg_handle *h) {
  const GraphImage &m = *h->img;
  const uint32_t nwarp = CELL_THREADS / 32, stages = 2;
  const uint32_t n_tab = (uint32_t)std::min<uint64_t>({m.nweight, (uint64_t)m.n_hot, (uint64_t)SELL_MAX_TAB});
  const size_t budget = 112 * 1024;  // two CTAs per SM
  const size_t fixed = sell_staged_ring_off(n_tab, stages, nwarp) + 128 + (size_t)nwarp * m.cell_kmax * 256;
  if (fixed + (size_t)nwarp * stages * 4 * 128 > budget) return fail(DWG_ERR_UNSUPPORTED, "categorical sweep: cardinality too large for shared memory");
  uint32_t cap = (uint32_t)((budget - fixed) / (nwarp * stages * 4)) & ~3u;
  if (const char *e = getenv("DWG_SELL_STAGE_WORDS")) cap = std::min<uint32_t>(cap, std::max(80, atoi(e)) & ~3u);  // tests
  const uint32_t words = std::max(80u, std::min(cap, (m.sell_max_words + 3u) & ~3u));
  const size_t smem = cell_smem(n_tab, stages, words, nwarp, m.cell_kmax);
  const bool allhot = m.nweight <= n_tab;
  int rc = allhot ? sell_grid(k_cell_infer_staged<true, CELL_THREADS>, smem, h->num_sms, &h->cell_blocks, CELL_THREADS)
                  : sell_grid(k_cell_infer_staged<false, CELL_THREADS>, smem, h->num_sms, &h->cell_blocks, CELL_THREADS);
  if (rc) return rc;
  h->cell_stages = stages;
  h->cell_stage_words = words;
  h->cg.sl_off = h->sl_off.p; h->cg.blob = h->blob.p;
  h->cg.nweight = (uint32_t)m.nweight; h->cg.n_tab = n_tab; h->cg.kmax = m.cell_kmax;
  return DWG_OK;
}
void cell_launch(dwg_handle *h, bool pdl, uint32_t blocks, const CellGraph &cg, const SweepParams &p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(CELL_THREADS);
  cfg.dynamicSmemBytes = cell_smem(cg.n_tab, h->cell_stages, h->cell_stage_words, CELL_THREADS / 32, cg.kmax);
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  if (h->img->nweight <= cg.n_tab)
    cudaLaunchKernelEx(&cfg, k_cell_infer_staged<true, CELL_THREADS>, cg, h->ss, p, h->cell_stages, h->cell_stage_words);
  else
    cudaLaunchKernelEx(&cfg, k_cell_infer_staged<false, CELL_THREADS>, cg, h->ss, p, h->cell_stages, h->cell_stage_words);
}
// k_cell_learn_live: the CTA's weight copy (28 B per hot weight), the potentials array and a
// two-stage ring per warp must fit two CTAs per SM; the hot set is the prefix of the internal
// (hottest first) order that leaves room for a ring stage holding the longest slice
int cell_learn_setup(dwg_handle *h) {
  const GraphImage &m = *h->img;
  const size_t budget = 112 * 1024, nwarp = SELL_LEARN_THREADS / 32;
  const size_t bufs = nwarp * (size_t)m.cell_kmax * 256 + 256 + nwarp * SELL_LEARN_STAGES * 8;
  const uint32_t want = std::max(80u, (m.sell_max_words + 3u) & ~3u);
  uint32_t words = want;
  if (const char *e = getenv("DWG_SELL_STAGE_WORDS")) words = std::min<uint32_t>(words, std::max(80, atoi(e)) & ~3u);  // tests
  const size_t ring = nwarp * SELL_LEARN_STAGES * 4 * (size_t)words;
  uint32_t n_hot = m.n_hot;
  if (bufs + ring + 28ull * n_hot > budget) {
    if (bufs + ring + 28ull * 256 <= budget) n_hot = (uint32_t)((budget - bufs - ring) / 28) & ~15u;
    else {  // long slices: a shorter stage (the rest of such a slice is read from HBM)
      n_hot = std::min<uint32_t>(n_hot, 256);
      if (bufs + 28ull * n_hot + nwarp * SELL_LEARN_STAGES * 4 * 80 > budget) return DWG_OK;  // no room: the CSR kernels learn
      words = std::max(80u, (uint32_t)((budget - bufs - 28ull * n_hot) / (nwarp * SELL_LEARN_STAGES * 4)) & ~3u);
    }
  }
  h->sg.sl_off = h->sl_off.p; h->sg.blob = h->blob.p;
  h->sg.nweight = (uint32_t)m.nweight; h->sg.n_hot = n_hot;
  h->learn_stage_words = words;
  int rc = sell_grid(k_cell_learn_live, cell_learn_smem(n_hot, words, m.cell_kmax), h->num_sms, &h->sell_blocks_learn,
                     SELL_LEARN_THREADS);
  if (rc) return rc;
  h->cell_learn = true;
  return DWG_OK;
}
void cell_learn_launch(dwg_handle *h, uint32_t blocks, uint32_t threads, bool live, const SweepParams &p, const SellLearn &L) {
  SellGraph sg = h->sg;
  if (live) sg.n_hot = 0;
  CellLearnExtra X{h->w_fixed.p, h->img->cell_kmax, h->opts.learn_non_evidence};
  k_cell_learn_live<<<blocks, threads, cell_learn_smem(sg.n_hot, h->learn_stage_words, X.kmax), h->stream>>>(
      sg, h->ss, p, L, X, h->learn_stage_words);
}
int sell_setup(dwg_handle *h) {
  const GraphImage &m = *h->img;
  int rc = DWG_OK;
#ifdef DWG_DEV_BUILD
  if (!(m.has_multi && m.fv_mode == FV_ONE && !m.has_wide))
    return fail(DWG_ERR_UNSUPPORTED, "development build (DWG_DEV_BUILD): only coupled Boolean graphs with unit feature values");
#endif
#define CALL(MU, FV, WD) rc = sell_setup_t<MU, FV, WD>(h)
  SELL_SWITCH(CALL)
#undef CALL
  return rc;
}
void sell_launch(dwg_handle *h, bool staged, bool pdl, uint32_t blocks, const SellGraph &sg, const SweepParams &p) {
  const GraphImage &m = *h->img;
#define CALL(MU, FV, WD) sell_launch_t<MU, FV, WD>(h, staged, pdl, blocks, sg, p)
  SELL_SWITCH(CALL)
#undef CALL
}
void sell_learn_launch(dwg_handle *h, uint32_t blocks, uint32_t threads, bool live, const SweepParams &p, const SellLearn &L) {
  const GraphImage &m = *h->img;
  if (m.cell) return cell_learn_launch(h, blocks, threads, live, p, L);
#define CALL(MU, FV, WD) sell_learn_launch_t<MU, FV, WD>(h, blocks, threads, live, p, L)
  SELL_SWITCH(CALL)
#undef CALL
}
}  // namespace

extern "C" {

const char *dwg_last_error(void) { return g_err.c_str(); }
const char *dwg_version(void) { return "dwg 0.1 (sm_100a)"; }
uint64_t dwg_launch_count(void) { return g_launches.load(); }

void dwg_opts_default(dwg_opts_t *o) {
  std::memset(o, 0, sizeof(*o));
  o->reg_param = 0.01;  // src/cmd_parser.cc:162
  o->schedule = DWG_SCHED_AUTO;
  o->seed = 0x5eed5eedull;
}

int dwg_graph_from_raw(const dwg_raw_t *raw, int schedule, dwg_graph **out) {
  if (!raw || !out) return fail(DWG_ERR_INVALID, "null argument");
  RawView v;
  v.nvar = raw->nvar; v.ndom = raw->ndom; v.nfac = raw->nfac; v.nweight = raw->nweight;
  v.var_role = raw->var_role; v.var_init = raw->var_init; v.var_dtype = raw->var_dtype;
  v.var_card = raw->var_card; v.dom_vid = raw->dom_vid; v.dom_ptr = raw->dom_ptr;
  v.dom_val = raw->dom_val; v.dom_truth = raw->dom_truth; v.fac_func = raw->fac_func;
  v.fac_ptr = raw->fac_ptr; v.edge_vid = raw->edge_vid; v.edge_eq = raw->edge_eq;
  v.fac_wid = raw->fac_wid; v.fac_fv = raw->fac_fv; v.w_init = raw->w_init; v.w_fixed = raw->w_fixed;
  static const uint64_t zero = 0;
  if (!v.fac_ptr) v.fac_ptr = &zero;
  if (!v.dom_ptr) v.dom_ptr = &zero;
  std::unique_ptr<dwg_graph> g(new dwg_graph);
  int err = 0;
  std::string msg;
  try {
    msg = compile_graph(v, schedule, &g->img, &err);
  } catch (const std::exception &e) {
    return fail(DWG_ERR_INVALID, std::string("compile: ") + e.what());
  }
  if (!msg.empty()) return fail(err, msg);
  *out = g.release();
  return DWG_OK;
}

int dwg_graph_from_files(const char *meta, int n_var, const char *const *var_files, int n_dom,
                         const char *const *dom_files, int n_fac, const char *const *fac_files,
                         int n_wgt, const char *const *wgt_files, int schedule, dwg_graph **out) {
  if (!meta || !out) return fail(DWG_ERR_INVALID, "null argument");
  try {
    MetaCounts m;
    std::string msg = read_meta(meta, &m);
    if (!msg.empty()) return fail(DWG_ERR_IO, msg);
    auto vec = [](int n, const char *const *p) {
      std::vector<std::string> v;
      for (int i = 0; i < n; ++i) v.emplace_back(p[i]);
      return v;
    };
    RawGraph raw;
    int err = 0;
    msg = load_raw(m, vec(n_var, var_files), vec(n_dom, dom_files), vec(n_fac, fac_files),
                   vec(n_wgt, wgt_files), &raw, &err);
    if (!msg.empty()) return fail(err, msg);
    std::unique_ptr<dwg_graph> g(new dwg_graph);
    msg = compile_graph(view_of(raw), schedule, &g->img, &err);
    if (!msg.empty()) return fail(err, msg);
    *out = g.release();
  } catch (const std::exception &e) {
    return fail(DWG_ERR_INVALID, std::string("load: ") + e.what());
  }
  return DWG_OK;
}

int dwg_graph_synthetic_unary(uint64_t nvar, uint32_t degree, uint64_t nweight, double zipf, int real_features,
                              double p_evid, uint64_t seed, double *cdf_out, dwg_graph **out) {
  if (!out) return fail(DWG_ERR_INVALID, "null argument");
  try {
    std::unique_ptr<dwg_graph> g(new dwg_graph);
    int err = 0;
    std::vector<double> cdf;
    const std::string msg = build_synthetic_unary(nvar, degree, nweight, zipf, real_features != 0, p_evid, seed, &g->img,
                                                  cdf_out ? &cdf : nullptr, &err);
    if (!msg.empty()) return fail(err, msg);
    if (cdf_out && !cdf.empty()) std::memcpy(cdf_out, cdf.data(), 8 * cdf.size());
    *out = g.release();
  } catch (const std::exception &e) {
    return fail(DWG_ERR_INVALID, std::string("synthetic graph: ") + e.what());
  }
  return DWG_OK;
}

void dwg_graph_free(dwg_graph *g) { delete g; }

int dwg_graph_set_page_locking(dwg_graph *g, int on) {
  if (!g) return fail(DWG_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(g->pin_mu);
  g->pin_wanted = on != 0;
  return DWG_OK;
}

static size_t image_device_bytes(const GraphImage &m) {
  size_t b = 0;
  b += m.var_flags.size() + m.w_fixed.size();
  b += (size_t)m.assign_bytes * m.nvar_pad * 3;  // evid_val + two chains
  if (m.sell || m.cell) b += 4 * (m.sl_off.size() + m.blob.size());
  if (m.cell) b += 4 * m.var_row.size();  // (the CSR stream joins when a learning sweep needs it)
  else if (!m.sell) {
    b += m.e_tag.size() + 4 * (m.var_row.size() + m.row_ptr.size() + m.row_oth.size() + m.e_wid.size() + m.oth.size());
    b += 4 * m.e_fv32.size() + 8 * m.e_fv64.size() + 8 * m.truthiness.size();
  }
  b += 4 * m.ndrow + 4 * m.row_of_value.size();  // tallies, their ABI-order map
  b += 8 * 3 * m.nweight;  // weights + prev + gradient sums
  b += 4 * 3 * (m.cls_N_evid.size());
  return b;
}

int dwg_graph_info(const dwg_graph *g, dwg_graph_info_t *o) {
  if (!g || !o) return fail(DWG_ERR_INVALID, "null argument");
  const GraphImage &m = g->img;
  std::memset(o, 0, sizeof(*o));
  o->num_variables = m.nvar; o->num_factors = m.nfac; o->num_weights = m.nweight;
  o->num_edges = m.nedge; o->num_values = m.nrow;
  o->num_variables_evidence = m.nvar_evid; o->num_variables_query = m.nvar_query;
  o->num_index_entries = m.nindex; o->num_other_refs = m.noth;
  o->num_colors = m.ncolor(); o->schedule = m.schedule;
  o->has_categorical = m.has_categorical; o->has_multi_arity = m.has_multi;
  o->fv_mode = m.fv_mode; o->assign_bytes = m.assign_bytes;
  o->device_bytes = image_device_bytes(m);
  o->bytes_per_sweep_infer = m.bytes_infer;
  o->bytes_per_sweep_infer_query = m.bytes_infer_query;
  o->bytes_per_sweep_learn = m.bytes_learn;
  o->num_device_rows = m.ndrow;
  return DWG_OK;
}

int dwg_graph_var_meta(const dwg_graph *g, uint8_t *dtype, uint8_t *is_evid, uint64_t *value_base,
                       uint64_t *values) {
  if (!g) return fail(DWG_ERR_INVALID, "null argument");
  const GraphImage &m = g->img;
  if (dtype) std::memcpy(dtype, m.var_dtype.data(), m.nvar);
  if (is_evid) std::memcpy(is_evid, m.var_is_evid.data(), m.nvar);
  if (value_base) {
    if (m.value_base.empty()) for (uint64_t v = 0; v <= m.nvar; ++v) value_base[v] = v;  // streamed Boolean image: one value per variable
    else std::memcpy(value_base, m.value_base.data(), 8 * (m.nvar + 1));
  }
  if (values) {
    if (m.values.empty()) std::memset(values, 0, 8 * m.nrow);
    else std::memcpy(values, m.values.data(), 8 * m.nrow);
  }
  return DWG_OK;
}

int dwg_debug_schedule(const dwg_graph *g, uint64_t *perm, uint64_t *color_start) {
  if (!g) return fail(DWG_ERR_INVALID, "null argument");
  const GraphImage &m = g->img;
  // pads (class fillers of sliced-ELL graphs) are not variables: report the compact order
  uint64_t k = 0;
  for (uint32_t c = 0; c < m.ncolor(); ++c) {
    if (color_start) color_start[c] = k;
    for (uint64_t d = m.color_start[c]; d < m.color_start[c + 1]; ++d)
      if (m.perm[d] != 0xffffffffu) {
        if (perm) perm[k] = m.perm[d];
        ++k;
      }
  }
  if (color_start) color_start[m.ncolor()] = k;
  return DWG_OK;
}

int dwg_debug_learn_order(const dwg_graph *g, uint64_t *perm, uint64_t *n_out) {
  if (!g) return fail(DWG_ERR_INVALID, "null argument");
  const GraphImage &m = g->img;
  uint64_t k = 0;
  if (m.sell || (m.cell && !m.learn_visit.empty() && !getenv("DWG_CELL_OFF"))) {
    for (uint32_t sl : m.learn_visit)
      for (uint64_t d = (uint64_t)sl * 32; d < (uint64_t)sl * 32 + 32; ++d)
        if (m.perm[d] != 0xffffffffu) {
          if (perm) perm[k] = m.perm[d];
          ++k;
        }
  } else {
    for (uint64_t d = 0; d < m.nvar_pad; ++d)
      if (m.perm[d] != 0xffffffffu) {
        if (perm) perm[k] = m.perm[d];
        ++k;
      }
  }
  if (n_out) *n_out = k;
  return DWG_OK;
}

int dwg_debug_image(const dwg_graph *g, const char *name, const void **data, uint64_t *count,
                    uint32_t *elem_size) {
  if (!g || !name || !data || !count || !elem_size) return fail(DWG_ERR_INVALID, "null argument");
  const GraphImage &m = g->img;
  const std::string n(name);
#define ARR(field)                                                    \
  if (n == #field) {                                                  \
    *data = m.field.data(); *count = m.field.size();                  \
    *elem_size = (uint32_t)sizeof(m.field[0]);                        \
    return DWG_OK;                                                    \
  }
  ARR(perm) ARR(inv) ARR(color_start) ARR(var_flags) ARR(evid_val) ARR(var_row) ARR(row_ptr)
  ARR(row_oth) ARR(truthiness) ARR(e_wid) ARR(e_fv32) ARR(e_fv64) ARR(e_tag) ARR(oth) ARR(w_init)
  ARR(w_fixed) ARR(row_of_value) ARR(values) ARR(value_base) ARR(w_perm) ARR(w_inv)
  ARR(sl_off) ARR(blob) ARR(learn_visit)
  ARR(cls_N_evid) ARR(cls_Q_evid) ARR(cls_N_all) ARR(cls_Q_all) ARR(cls_H_evid) ARR(cls_H_all)
#undef ARR
  if (n == "oth_packed") {
    static thread_local uint32_t v;
    v = m.oth_packed; *data = &v; *count = 1; *elem_size = 4;
    return DWG_OK;
  }
  return fail(DWG_ERR_INVALID, "unknown image array " + n);
}

// ---------------------------------------------------------------------------------
static int init_state(dwg_handle *h) {
  // InferenceResult ctor, src/inference_result.cc:24-42: both chains start at the evidence
  // value (0 for query variables); weights from the Weight records; tallies cleared.
  const GraphImage &m = *h->img;
  if (m.assign_bytes == 1) {
    CU(cudaMemcpyAsync(h->assign_free8.p, h->evid_val8.p, m.nvar_pad, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(h->assign_evid8.p, h->evid_val8.p, m.nvar_pad, cudaMemcpyDeviceToDevice, h->stream));
  } else {
    CU(cudaMemcpyAsync(h->assign_free32.p, h->evid_val32.p, 4 * m.nvar_pad, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(h->assign_evid32.p, h->evid_val32.p, 4 * m.nvar_pad, cudaMemcpyDeviceToDevice, h->stream));
  }
  CU(cudaMemsetAsync(h->tallies.p, 0, 4 * m.ndrow + 16, h->stream));
  CU(cudaMemsetAsync(h->grad_acc.p, 0, 8 * m.nweight + 16, h->stream));
  h->infer_sweeps = 0;
  return DWG_OK;
}

// CSR stream on the device (graph_image.h): resident from the start for categorical graphs,
// uploaded on first use for sliced-ELL graphs
static int ensure_csr(dwg_handle *h) {
  if (h->csr_ready) return DWG_OK;
  const GraphImage &m = *h->img;
  if (!m.csr_built)
    return fail(DWG_ERR_UNSUPPORTED, "a streamed synthetic graph has no CSR stream (parity hooks and the sequential schedule need one)");
  cudaStream_t s = h->stream;
  CU(cudaSetDevice(h->device));
  CU(h->e_tag.upload(m.e_tag, s));
  CU(h->row_ptr.upload(m.row_ptr, s));
  CU(h->e_wid.upload(m.e_wid, s));
  if (m.has_multi) { CU(h->row_oth.upload(m.row_oth, s)); CU(h->oth.upload(m.oth, s)); }
  if (m.fv_mode == FV_F32) CU(h->e_fv32.upload(m.e_fv32, s));
  if (m.fv_mode == FV_F64) CU(h->e_fv64.upload(m.e_fv64, s));
  if (m.has_truthiness) CU(h->truthiness.upload(m.truthiness, s));
  CU(cudaStreamSynchronize(s));
  DevGraph &dg = h->dg;
  dg.row_ptr = h->row_ptr.p;
  dg.row_oth = m.has_multi ? h->row_oth.p : nullptr;
  dg.truthiness = m.has_truthiness ? h->truthiness.p : nullptr;
  dg.e_wid = h->e_wid.p;
  dg.e_fv32 = m.fv_mode == FV_F32 ? h->e_fv32.p : nullptr;
  dg.e_fv64 = m.fv_mode == FV_F64 ? h->e_fv64.p : nullptr;
  dg.e_tag = h->e_tag.p;
  dg.oth = h->oth.p;
  h->csr_ready = true;
  return DWG_OK;
}

int dwg_create(const dwg_graph *g, const dwg_opts_t *opts, int device, dwg_handle **out) {
  if (!g || !out) return fail(DWG_ERR_INVALID, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(DWG_ERR_CUDA, "no CUDA device: the Gibbs sweep has no CPU path");
  if (device < 0 || device >= ndev) return fail(DWG_ERR_CUDA, "bad device index");
  CU(cudaSetDevice(device));
  {
    double tab[DWG_LOG2_TABLE];
    for (int k = 0; k < DWG_LOG2_TABLE; ++k) tab[k] = std::log2(1.0 + (double)k);
    CU(cudaMemcpyToSymbol(c_log2_1p, tab, sizeof(tab)));
  }
  std::unique_ptr<dwg_handle> h(new dwg_handle);
  h->img = &g->img;
  if (opts) h->opts = *opts; else dwg_opts_default(&h->opts);
  h->device = device;
  CU(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  const GraphImage &m = g->img;
  cudaStream_t s = h->stream;
  const bool timing = getenv("DWG_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    auto n = std::chrono::steady_clock::now();
    if (timing) fprintf(stderr, "[create] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t_prev).count());
    t_prev = n;
  };
  const bool defer = h->opts.defer_upload != 0;
  if (!defer) g->pin();
  lap("page-lock (first create only)");
  struct ScopeGuard {
    ScopeGuard(std::vector<std::pair<void *, size_t>> *seg, bool d) { g_upload.segments = seg; g_upload.defer = d; }
    ~ScopeGuard() { g_upload = UploadScope(); }
  } scope_guard(&h->segments, defer);
  if (defer && (!m.sell && !m.cell)) return fail(DWG_ERR_UNSUPPORTED, "defer_upload: sliced-ELL graphs only");
  CU(h->var_flags.upload(m.var_flags, s));
  if (m.has_categorical) CU(h->var_row.upload(m.var_row, s));
  std::vector<uint8_t> fixed_dev(m.nweight);
  std::vector<double> w_dev(m.nweight);
  for (uint64_t i = 0; i < m.nweight; ++i) {
    w_dev[i] = m.w_init[m.w_perm[i]];
    fixed_dev[i] = m.w_fixed[m.w_perm[i]];
  }
  if (m.assign_bytes == 1) {
    CU(h->evid_val8.upload(m.evid_val8, s));
    CU(h->assign_free8.alloc(m.nvar_pad));
    CU(h->assign_evid8.alloc(m.nvar_pad));
  } else {
    CU(h->evid_val32.upload(m.evid_val, s));
    CU(h->assign_free32.alloc(m.nvar_pad));
    CU(h->assign_evid32.alloc(m.nvar_pad));
  }
  CU(h->tallies.alloc(m.ndrow));
  // dwg_get_tallies: schedule order -> ABI order on the device.  Allocated here, copied behind
  // everything else on the `late` stream (below).  It is not one of the segments a peer receives:
  // only the replica that hands tallies to the host reads it, and a deferred replica that does
  // uploads its own copy on first use (get_tallies_impl).
  CU(h->row_of_value.alloc(m.row_of_value.size()));
  CU(h->w_fixed.upload(fixed_dev, s));
  CU(h->weights.upload(w_dev, s));
  CU(h->prev_weights.upload(w_dev, s));
  CU(h->grad_acc.alloc(m.nweight));
  CU(h->norms.alloc(2));
  // which variables take SGD steps decides the per-launch update counts (compile.cc)
  {
    const bool all = h->opts.learn_non_evidence;
    CU(h->cls_N.upload(all ? m.cls_N_all : m.cls_N_evid, s));
    CU(h->cls_Q.upload(all ? m.cls_Q_all : m.cls_Q_evid, s));
    CU(h->cls_H.upload(all ? m.cls_H_all : m.cls_H_evid, s));
  }
  // whole-graph per-weight statistics of a learning sweep (categorical sliced-ELL graphs keep
  // per-class rows for the CSR kernels: summed here)
  std::vector<float> lvN, lvH;
  if (m.cell) {
    const bool all = h->opts.learn_non_evidence;
    const std::vector<float> &N = all ? m.cls_N_all : m.cls_N_evid, &H = all ? m.cls_H_all : m.cls_H_evid;
    lvN.assign(m.nweight, 0.f); lvH.assign(m.nweight, 0.f);
    const uint64_t rows = m.nweight ? N.size() / m.nweight : 0;
    for (uint64_t r = 0; r < rows; ++r)
      for (uint64_t i = 0; i < m.nweight; ++i) { lvN[i] += N[r * m.nweight + i]; lvH[i] += H[r * m.nweight + i]; }
    if (m.cls_shared)  // one class-averaged row
      for (uint64_t i = 0; i < m.nweight; ++i) { lvN[i] *= m.ncolor(); lvH[i] *= m.ncolor(); }
    CU(h->lv_N.upload(lvN, s));
    CU(h->lv_H.upload(lvH, s));
    CU(h->learn_visit.upload(m.learn_visit, s));
  }
  if (m.sell || m.cell) {
    if (m.sell) {
    CU(h->sl_off.upload(m.sl_off, s));
    CU(h->blob.upload(m.blob, s));
    CU(h->learn_visit.upload(m.learn_visit, s));
    }
    if (!m.part_slices.empty()) {
      const size_t nsl = m.sl_off.size() - 1;
      CU(h->part_sums.alloc(4 * nsl));
      CU(h->part_done.alloc(nsl));
      CU(cudaMemsetAsync(h->part_done.p, 0, 4 * nsl, s));
    }
    // Boolean variables have no truthiness: under --noise_aware only --learn_non_evidence learns
    h->learns = h->opts.learn_non_evidence || !h->opts.noise_aware;
    // the epoch plan of launch_learn_sell: curvature bounds of the learnable weights, sorted, and
    // the coupling factor (SGD steps of one variable move all its weights together)
    const bool all = h->opts.learn_non_evidence;
    const std::vector<float> &N = m.cell ? lvN : all ? m.cls_N_all : m.cls_N_evid, &H = m.cell ? lvH : all ? m.cls_H_all : m.cls_H_evid;
    double updates = 0;
    for (uint64_t i = 0; i < m.nweight; ++i) {
      updates += N[i];
      if (H[i] > 0) h->learn_H_sorted.push_back(H[i]);
    }
    std::sort(h->learn_H_sorted.begin(), h->learn_H_sorted.end());
    const double learners = (double)std::max<uint64_t>(1, all ? m.nvar : m.nvar_evid);
    h->learn_rho = std::max(1.0, updates / learners / 10.0);
  }
  lap("allocations + copies queued");
  int rc = defer ? DWG_OK : init_state(h.get());  // (deferred: dwg_reset_state once the graph has arrived)
  if (rc) return rc;
  // (no synchronisation yet: the kernel setup below is host work -- occupancy queries, function
  // attributes, the epoch plan -- and runs while the copies are on the link)
  h->device_bytes = image_device_bytes(m);

  DevGraph &dg = h->dg;
  dg.nvar = (uint32_t)m.nvar_pad; dg.nrow = (uint32_t)m.ndrow;
  dg.var_flags = h->var_flags.p;
  dg.evid_val = m.assign_bytes == 1 ? (const void *)h->evid_val8.p : (const void *)h->evid_val32.p;
  dg.var_row = m.has_categorical ? h->var_row.p : nullptr;
  dg.oth_packed = m.oth_packed;
  dg.w_fixed = h->w_fixed.p;
  DevState &ds = h->ds;
  ds.assign_free = m.assign_bytes == 1 ? (void *)h->assign_free8.p : (void *)h->assign_free32.p;
  ds.assign_evid = m.assign_bytes == 1 ? (void *)h->assign_evid8.p : (void *)h->assign_evid32.p;
  ds.weights = h->weights.p;
  ds.grad_acc = h->grad_acc.p;
  ds.tallies = h->tallies.p;
  if (m.cell) {
    // categorical sliced-ELL graph: the inference sweep streams the blob; the CSR stream (learning
    // sweeps, parity hooks) is uploaded on first use
    CU(h->sl_off.upload(m.sl_off, s));
    CU(h->blob.upload(m.blob, s));
    SellState &ss = h->ss;
    ss.assign_free = h->assign_free8.p; ss.assign_evid = h->assign_evid8.p;
    ss.weights = h->weights.p; ss.grad_acc = h->grad_acc.p; ss.tallies = h->tallies.p;
    rc = cell_setup(h.get());
    if (rc) return rc;
    rc = cell_learn_setup(h.get());
    if (rc) return rc;
    h->cls_query_end = g->query_end();
  } else if (!m.sell) {
    rc = ensure_csr(h.get());
    if (rc) return rc;
  } else {
    SellGraph &sg = h->sg;
    sg.sl_off = h->sl_off.p; sg.blob = h->blob.p;
    sg.nweight = (uint32_t)m.nweight; sg.n_hot = m.n_hot;
    SellState &ss = h->ss;
    ss.assign_free = h->assign_free8.p; ss.assign_evid = h->assign_evid8.p;
    ss.weights = h->weights.p; ss.grad_acc = h->grad_acc.p; ss.tallies = h->tallies.p;
    ss.part_sums = h->part_sums.p; ss.part_done = h->part_done.p;
    rc = sell_setup(h.get());
    if (rc) return rc;
    h->cls_query_end = g->query_end();
  }
  lap("kernel setup");
  CU(cudaStreamSynchronize(s));  // the staging vectors go out of scope
  lap("copies done");
  if (!defer && !m.row_of_value.empty()) {
    // (stream s has drained: this copy has the link to itself and overlaps whatever the caller does next)
    CU(cudaStreamCreateWithFlags(&h->late, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->late_done, cudaEventDisableTiming));
    CU(cudaMemcpyAsync(h->row_of_value.p, m.row_of_value.data(), 4 * m.row_of_value.size(), cudaMemcpyHostToDevice, h->late));
    CU(cudaEventRecord(h->late_done, h->late));
    h->row_map_ready = true;
  }
  *out = h.release();
  return DWG_OK;
}

int dwg_release_cached_memory(int device) {
  CU(cudaSetDevice(device));
  if (!dev_pool_on(device)) return DWG_OK;
  cudaMemPool_t pool = nullptr;
  CU(cudaDeviceGetDefaultMemPool(&pool, device));
  CU(cudaStreamSynchronize((cudaStream_t)0));
  CU(cudaMemPoolTrimTo(pool, 0));
  return DWG_OK;
}

void dwg_destroy(dwg_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  delete h;
}

int dwg_graph_segments(dwg_handle *h, int max, void **ptrs, uint64_t *bytes, int *n) {
  if (!h || !n) return fail(DWG_ERR_INVALID, "null argument");
  *n = (int)h->segments.size();
  for (int i = 0; i < *n && i < max; ++i) {
    if (ptrs) ptrs[i] = h->segments[i].first;
    if (bytes) bytes[i] = h->segments[i].second;
  }
  return DWG_OK;
}

int dwg_reset_state(dwg_handle *h) {
  if (!h) return fail(DWG_ERR_INVALID, "null handle");
  h->chain_open = false;
  CU(cudaSetDevice(h->device));
  return init_state(h);
}

static SweepParams make_params(dwg_handle *h, double stepsize) {
  SweepParams p{};
  const uint64_t key = h->opts.seed;
  p.key0 = (uint32_t)key;
  p.key1 = (uint32_t)(key >> 32) ^ (h->opts.replica * 0x9E3779B9u + h->opts.replica);
  p.sweep_lo = (uint32_t)h->sweep_counter;
  p.sweep_hi = (uint32_t)(h->sweep_counter >> 32);
  p.sample_evidence = h->opts.sample_evidence;
  p.learn_non_evidence = h->opts.learn_non_evidence;
  p.noise_aware = h->opts.noise_aware;
  p.reg_l1 = h->opts.regularization == 1;
  p.reg_param = h->opts.reg_param;
  p.stepsize = stepsize;
  p.sequential = h->sequential || h->img->schedule == DWG_SCHED_SEQUENTIAL;
  p.r_free = h->explicit_r ? h->r_free.p : nullptr;
  p.r_evid = h->explicit_r ? h->r_evid.p : nullptr;
  p.unit_stride = 1;
  p.unit_phase = 0;
  return p;
}

static double env_double(const char *name, double dflt) {
  const char *e = getenv(name);
  return e && *e ? atof(e) : dflt;
}

// One learning sweep of a sliced-ELL graph (k_sell_learn_live, dwg_sell.cuh): the plan of the epoch.
//   H_k   curvature bound of weight k over a sweep (sum of 1/4 (|feature| x sign range)^2 over its
//         learnable records): stepsize x H_k = relaxation times weight k goes through per sweep
//   x_k   = stepsize x H_k x rho x (slices in flight / slices): step mass in flight; a weight keeps
//         the reference's step while x_k <= cap, beyond that its step is scaled by cap / x_k
// The tail (the last c relaxation times of the median fast weight, at least 12 rounds) runs on as
// few warps as keep the 90th-percentile weight at x <= cap with s = 1; the head runs on every SM.
// A sweep short enough for the narrow grid (<= rounds_max rounds) is all tail.
static int launch_learn_sell(dwg_handle *h, SweepParams &p, double stepsize) {
  const GraphImage &m = *h->img;
  const uint32_t nslice = (uint32_t)m.learn_visit.size();
  if (!nslice) return DWG_OK;
  const double cap = env_double("DWG_LEARN_XCAP", 0.15), tail_c = env_double("DWG_LEARN_TAILC", 3.0);
  const uint32_t rounds_max = (uint32_t)env_double("DWG_LEARN_ROUNDS", 96), period_wide = std::max(1u, (uint32_t)env_double("DWG_LEARN_PERIOD", 32));
  const uint32_t max_warps = (uint32_t)h->sell_blocks_learn * (SELL_LEARN_THREADS / 32);
  const std::vector<float> &Hs = h->learn_H_sorted;
  const double a = stepsize > 0 ? stepsize : 0.0, rho = h->learn_rho;
  // warps that keep weight curvature H at x <= cap with one slice per warp in flight
  auto narrow_warps = [&](double H) -> uint32_t {
    if (!(a * H * rho > 0)) return max_warps;
    return (uint32_t)std::max(1.0, std::min((double)max_warps, std::floor(cap * nslice / (a * H * rho))));
  };
  // A learning sweep is Hogwild across colour classes (like the reference's threads): the slices in
  // flight must stay a small part of the graph, or neighbouring variables are resampled in lockstep
  // from each other's old values (synchronous Gibbs: on a strongly coupled graph the free chain
  // then never settles).  At most 1/16 of the slices on the narrow grid, 1/4 on the wide one.
  const uint32_t narrow_max = std::max(1u, nslice / 16), wide_warps = std::max(1u, std::min(max_warps, nslice / 4));
  uint32_t tail_slices = 0, tail_warps = std::min(max_warps, narrow_max);
  if (const char *e = getenv("DWG_LEARN_WARPS")) {  // tests: 1 = one warp walks the visit order slice by slice
    tail_slices = nslice;
    tail_warps = (uint32_t)std::max(1, atoi(e));
  } else if (h->learns && !Hs.empty() && a > 0) {
    const size_t n = Hs.size();
    const uint32_t g_all = std::min(narrow_max, narrow_warps(Hs[(size_t)(0.9 * (n - 1))]));
    if ((nslice + g_all - 1) / g_all <= rounds_max) {
      tail_slices = nslice; tail_warps = g_all;  // short sweep: all of it at the reference's step
    } else {
      const size_t lo = std::lower_bound(Hs.begin(), Hs.end(), (float)(1.0 / a)) - Hs.begin();  // fast: stepsize x H > 1
      if (lo < n) {
        const double h_med = Hs[lo + (n - lo) / 2], h_ref = Hs[lo + (size_t)(0.9 * (n - 1 - lo))];
        tail_warps = std::min(narrow_max, narrow_warps(h_ref));
        const double f = std::min(1.0, tail_c / (a * h_med));
        tail_slices = (uint32_t)std::min<double>(nslice, std::max(std::ceil(f * nslice), 12.0 * tail_warps));
        tail_warps = std::max(1u, std::min(tail_warps, tail_slices / 12u));
      }
    }
  }
  SellLearn L{};
  L.visit = h->learn_visit.p;
  L.step = a;
  L.decay = !h->learns ? 0.0 : p.reg_l1 ? h->opts.reg_param : h->opts.reg_param * a;
  L.inv_nslice = 1.0 / nslice;
  L.reg_l1 = p.reg_l1;
  L.w_N = m.cell ? h->lv_N.p : h->cls_N.p; L.w_H = m.cell ? h->lv_H.p : h->cls_H.p;
  L.weights = h->weights.p;
  // period 0: the narrow form -- one warp per CTA (an SM to itself while there are SMs), weights
  // read from and pushed to the live vector record by record
  auto launch = [&](uint32_t vb, uint32_t ve, uint32_t warps, uint32_t period) {
    if (ve <= vb) return;
    const bool live = period == 0;
    // warps per CTA: eight on the wide grid; the narrow tail (period 1: a refresh and a flush per
    // slice) spreads its warps over more SMs, DWG_LEARN_TAIL_WPC per CTA
    const uint32_t tail_wpc = (uint32_t)std::max(1.0, std::min(8.0, env_double("DWG_LEARN_TAIL_WPC", 8)));
    const uint32_t wpc = live ? 1 : (vb != 0 || ve != nslice || period == 1) && period == 1 ? tail_wpc : SELL_LEARN_THREADS / 32;
    warps = std::max(1u, std::min(warps, ve - vb));
    uint32_t threads = wpc * 32, blocks = warps / wpc;
    if (warps < wpc) { threads = warps * 32; blocks = 1; }
    const uint32_t grid_warps = blocks * (threads / 32);
    L.vbegin = vb; L.vend = ve;
    L.period = live ? 0xffffffffu : period;
    const double inflight = std::min(1.0, (double)grid_warps * (live ? 1 : period) / nslice);
    // one warp cannot be narrowed further: it keeps the reference's step whatever x is
    L.xcap_inv = grid_warps == 1 ? 0.0 : a * rho * inflight / cap;
    L.visit = h->learn_visit.p;
    L.inv_nslice = 1.0 / nslice;
    sell_learn_launch(h, blocks, threads, live, p, L);
    ++g_launches;
    // the decay of the weights outside the CTAs' shared copies for this launch's share of the sweep
    const uint32_t n_hot = live ? 0u : h->sg.n_hot;
    if (m.nweight > n_hot && L.decay != 0.0) {
      const uint32_t ncold = (uint32_t)m.nweight - n_hot;
      k_cold_decay<<<(ncold + 255) / 256, 256, 0, h->stream>>>(L, h->w_fixed.p, n_hot, (uint32_t)m.nweight, (double)(ve - vb) / nslice);
      ++g_launches;
    }
  };
  launch(0, nslice - tail_slices, wide_warps, period_wide);
  // tail: the CTA's weight copy is refreshed before every slice (period 1: eight warps read the
  // live vector together, sweep a slice each, push their steps together).  Measured on c2_small
  // (scripts/learn_fidelity.py, 32 seeds): mean |dw| 0.0114 against a noise floor of 0.0101; the
  // free-running per-record form (period 0) gives 0.018 at the same width.
  // (a tail already as wide as the GPU has room left under the cap: it goes into the period)
  uint32_t tail_period = 1;
  if (tail_warps >= max_warps && !Hs.empty() && a > 0) {
    const size_t n = Hs.size(), lo = std::lower_bound(Hs.begin(), Hs.end(), (float)(1.0 / a)) - Hs.begin();
    const double h_ref = lo < n ? Hs[lo + (size_t)(0.9 * (n - 1 - lo))] : Hs[(size_t)(0.9 * (n - 1))];
    tail_period = (uint32_t)std::max(1.0, std::min((double)period_wide, std::floor(cap * nslice / (a * h_ref * rho * max_warps))));
  }
  launch(nslice - tail_slices, nslice, tail_warps, (uint32_t)env_double("DWG_LEARN_TAIL_PERIOD", tail_period));
  if (getenv("DWG_TIMING"))
    fprintf(stderr, "[learn] step %.4g: %u slices, head %u on %u warps (period %u), tail %u on %u warps (period %u)\n", a, nslice,
            nslice - tail_slices, wide_warps, period_wide, tail_slices, tail_warps, tail_period);
  return DWG_OK;
}

static int launch_sweep(dwg_handle *h, bool learn, double stepsize) {
  const GraphImage &m = *h->img;
  CU(cudaSetDevice(h->device));
  SweepParams p = make_params(h, stepsize);
  ++h->sweep_counter;
  const bool use_sell = m.sell && !p.sequential;
  const bool use_cell = m.cell && !p.sequential && !learn && !getenv("DWG_CELL_OFF");  // categorical inference sweeps
  // categorical sliced-ELL graphs learn through k_cell_learn_live (hard evidence; --noise_aware
  // steps every value by its truthiness: the CSR kernels)
  const bool cell_learn = m.cell && learn && !p.sequential && h->cell_learn && !p.noise_aware && !getenv("DWG_CELL_OFF");
  if (!use_sell && !use_cell && !cell_learn) {
    h->chain_open = false;
    int rc = ensure_csr(h);
    if (rc) return rc;
  }
  if ((use_sell && learn) || cell_learn) {
    h->chain_open = false;
    int rc = launch_learn_sell(h, p, stepsize);
    if (rc) return rc;
    CU(cudaGetLastError());
    return DWG_OK;
  }
  // CSR (categorical) learning: every class is cut into `chunks` CONTIGUOUS parts (ascending
  // position = ascending id block inside the evidence part) and the sweep runs part 0 of every
  // class, then part 1 of every class, ...: variables are visited in ascending id block like the
  // reference's scan (the learned weights depend on it) while every launch stays
  // inside one colour class.  A batch weight update (k_apply_weights) follows each launch.
  uint32_t chunks = 1;
  if (learn && !p.sequential && !use_sell) {
    chunks = h->opts.learn_chunks > 0 ? (uint32_t)h->opts.learn_chunks
                                      : std::max(1u, (uint32_t)std::llround(32.0 / (double)std::max(1u, m.ncolor())));
  }
  for (uint32_t b = 0; b < chunks; ++b)
  for (uint32_t c = 0; c < m.ncolor(); ++c) {
    const uint32_t cb = m.color_start[c], ce = m.color_start[c + 1];
    if (ce <= cb) continue;
    const uint32_t units = (ce - cb + 31) / 32;  // 32-variable units (= slices of the sliced-ELL streams)
    const uint32_t u0 = (uint32_t)((uint64_t)units * b / chunks), u1 = (uint32_t)((uint64_t)units * (b + 1) / chunks);
    if (u1 <= u0) continue;
    p.d_begin = cb + u0 * 32;
    p.d_end = std::min(ce, cb + u1 * 32);
    p.unit_stride = 1;
    p.unit_phase = 0;
    const uint32_t my_units = u1 - u0;
    {
      if (use_cell) {
        CellGraph cg = h->cg;
        cg.slice_begin = p.d_begin / 32;
        cg.slice_end = p.sample_evidence ? p.d_end / 32 : h->cls_query_end[c];  // gibbs_sampler.h:157
        if (cg.slice_end <= cg.slice_begin) continue;
        const uint32_t need = (cg.slice_end - cg.slice_begin + CELL_THREADS / 32 - 1) / (CELL_THREADS / 32);
        const bool pdl = !getenv("DWG_SELL_NO_PDL") && h->chain_open;
        cell_launch(h, pdl, std::min<uint32_t>(need, (uint32_t)h->cell_blocks), cg, p);
        h->chain_open = true;
      } else if (use_sell) {
        SellGraph sg = h->sg;
        sg.slice_begin = p.d_begin / 32;
        sg.slice_end = p.d_end / 32;
        sg.slice_stride = 1;
        if (!p.sample_evidence) sg.slice_end = h->cls_query_end[c];  // gibbs_sampler.h:157: the evidence tail of the class is not read at all
        if (sg.slice_end <= sg.slice_begin) continue;
        const uint32_t n_sl = sg.slice_end - sg.slice_begin;
        const bool staged = h->sell_stages != 0;
        const uint32_t wpb = staged ? h->sell_staged_warps : SELL_THREADS / 32;
        const uint32_t need = (n_sl + wpb - 1) / wpb;
        const uint32_t cap = (uint32_t)(staged ? h->sell_blocks_staged : h->sell_blocks_infer);
        // programmatic dependent launch behind the previous colour class of an inference run
        const bool pdl = !getenv("DWG_SELL_NO_PDL") && staged && h->chain_open;
        sell_launch(h, staged, pdl, std::min<uint32_t>(need, cap), sg, p);
        h->chain_open = staged;
      } else {
        uint32_t blocks = (my_units * 32 + SWEEP_THREADS - 1) / SWEEP_THREADS;
        uint32_t threads = SWEEP_THREADS;
        if (p.sequential) { blocks = 1; threads = 1; }
        if (m.assign_bytes == 1) {
          if (learn) k_sweep_learn<uint8_t><<<blocks, threads, 0, h->stream>>>(h->dg, h->ds, p);
          else k_sweep_infer<uint8_t><<<blocks, threads, 0, h->stream>>>(h->dg, h->ds, p);
        } else {
          if (learn) k_sweep_learn<uint32_t><<<blocks, threads, 0, h->stream>>>(h->dg, h->ds, p);
          else k_sweep_infer<uint32_t><<<blocks, threads, 0, h->stream>>>(h->dg, h->ds, p);
        }
      }
      ++g_launches;
      if (learn && !p.sequential && m.nweight) {
        // integrate this launch's SGD steps (k_apply_weights): the CSR kernel sums gradient + decay
        const size_t row = m.cls_shared ? 0 : (size_t)c * m.nweight;
        k_apply_weights<<<(uint32_t)((m.nweight + 255) / 256), 256, 0, h->stream>>>(
            h->weights.p, h->grad_acc.p, (const float *)(h->cls_N.p + row), (const float *)(h->cls_Q.p + row),
            (const float *)(h->cls_H.p + row), (const uint8_t *)h->w_fixed.p, (uint32_t)m.nweight, stepsize,
            h->opts.reg_param, (int)p.reg_l1, 1, (double)my_units / (double)units);
        ++g_launches;
      }
    }
  }
  CU(cudaGetLastError());
  return DWG_OK;
}

int dwg_learn_epoch(dwg_handle *h, double stepsize) {
  if (!h) return fail(DWG_ERR_INVALID, "null handle");
  return launch_sweep(h, true, stepsize);
}

int dwg_infer_epoch(dwg_handle *h) {
  if (!h) return fail(DWG_ERR_INVALID, "null handle");
  int rc = launch_sweep(h, false, 0.0);
  if (rc == DWG_OK) ++h->infer_sweeps;
  return rc;
}

// n epochs of the sequential schedule in one launch (k_seq_epochs)
static int launch_seq_epochs(dwg_handle *h, uint64_t n, bool learn, double stepsize, double decay) {
  const GraphImage &m = *h->img;
  CU(cudaSetDevice(h->device));
  h->chain_open = false;
  int rc = ensure_csr(h);
  if (rc) return rc;
  while (n) {
    const uint32_t chunk = (uint32_t)std::min<uint64_t>(n, 1u << 20);
    SweepParams p = make_params(h, stepsize);
    p.d_begin = 0; p.d_end = (uint32_t)m.nvar_pad;
    if (m.assign_bytes == 1) k_seq_epochs<uint8_t><<<1, 1, 0, h->stream>>>(h->dg, h->ds, p, chunk, learn ? 1 : 0, decay);
    else k_seq_epochs<uint32_t><<<1, 1, 0, h->stream>>>(h->dg, h->ds, p, chunk, learn ? 1 : 0, decay);
    ++g_launches;
    h->sweep_counter += chunk;
    if (!learn) h->infer_sweeps += chunk;
    for (uint32_t e = 0; e < chunk && learn; ++e) stepsize *= decay;
    n -= chunk;
  }
  CU(cudaGetLastError());
  return DWG_OK;
}
static bool is_sequential(const dwg_handle *h) { return h->sequential || h->img->schedule == DWG_SCHED_SEQUENTIAL; }

int dwg_infer_epochs(dwg_handle *h, uint64_t n) {
  if (!h) return fail(DWG_ERR_INVALID, "null handle");
  if (is_sequential(h) && n > 1 && !h->explicit_r) return launch_seq_epochs(h, n, false, 0.0, 1.0);
  for (uint64_t i = 0; i < n; ++i) {
    int rc = dwg_infer_epoch(h);
    if (rc) return rc;
  }
  return DWG_OK;
}

int dwg_learn_epochs(dwg_handle *h, uint64_t n, double stepsize, double decay) {
  if (!h) return fail(DWG_ERR_INVALID, "null handle");
  if (is_sequential(h) && n > 1 && !h->explicit_r) return launch_seq_epochs(h, n, true, stepsize, decay);
  for (uint64_t i = 0; i < n; ++i) {
    int rc = dwg_learn_epoch(h, stepsize);
    if (rc) return rc;
    stepsize *= decay;
  }
  return DWG_OK;
}

int dwg_sync(dwg_handle *h) {
  if (!h) return fail(DWG_ERR_INVALID, "null handle");
  h->chain_open = false;
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  return DWG_OK;
}

void *dwg_stream(dwg_handle *h) {
  if (h) h->chain_open = false;  // the caller may enqueue anything behind our launches
  return h ? (void *)h->stream : nullptr;
}
void *dwg_weights_device_ptr(dwg_handle *h) { return h ? (void *)h->weights.p : nullptr; }
void *dwg_tallies_device_ptr(dwg_handle *h) { return h ? (void *)h->tallies.p : nullptr; }

int dwg_tallies_pack(dwg_handle *h, int fields, void **words, uint64_t *nwords) {
  if (!h || !words || !nwords) return fail(DWG_ERR_INVALID, "null argument");
  if (fields != 1 && fields != 2 && fields != 4) return fail(DWG_ERR_INVALID, "fields must be 1, 2 or 4");
  h->chain_open = false;
  CU(cudaSetDevice(h->device));
  const uint64_t n = h->img->ndrow, nw = (n + fields - 1) / fields;
  if (h->tallies_packed.n < nw) {
    h->tallies_packed.reset();
    CU(h->tallies_packed.alloc(nw));
  }
  const uint32_t blocks = (uint32_t)((nw + 255) / 256);
  if (nw) {
    if (fields == 4) k_pack_tallies<4><<<blocks, 256, 0, h->stream>>>(h->tallies_packed.p, h->tallies.p, (uint32_t)n);
    else if (fields == 2) k_pack_tallies<2><<<blocks, 256, 0, h->stream>>>(h->tallies_packed.p, h->tallies.p, (uint32_t)n);
    else k_pack_tallies<1><<<blocks, 256, 0, h->stream>>>(h->tallies_packed.p, h->tallies.p, (uint32_t)n);
    ++g_launches;
  }
  CU(cudaGetLastError());
  *words = h->tallies_packed.p;
  *nwords = nw;
  return DWG_OK;
}

int dwg_tallies_unpack(dwg_handle *h, int fields) {
  if (!h) return fail(DWG_ERR_INVALID, "null handle");
  if (fields != 1 && fields != 2 && fields != 4) return fail(DWG_ERR_INVALID, "fields must be 1, 2 or 4");
  const uint64_t n = h->img->ndrow, nw = (n + fields - 1) / fields;
  if (h->tallies_packed.n < nw || !h->tallies_packed.p) return fail(DWG_ERR_INVALID, "dwg_tallies_unpack without dwg_tallies_pack");
  h->chain_open = false;
  CU(cudaSetDevice(h->device));
  const uint32_t blocks = (uint32_t)((nw + 255) / 256);
  if (nw) {
    if (fields == 4) k_unpack_tallies<4><<<blocks, 256, 0, h->stream>>>(h->tallies.p, h->tallies_packed.p, (uint32_t)n);
    else if (fields == 2) k_unpack_tallies<2><<<blocks, 256, 0, h->stream>>>(h->tallies.p, h->tallies_packed.p, (uint32_t)n);
    else k_unpack_tallies<1><<<blocks, 256, 0, h->stream>>>(h->tallies.p, h->tallies_packed.p, (uint32_t)n);
    ++g_launches;
  }
  CU(cudaGetLastError());
  return DWG_OK;
}

int dwg_clear_tallies(dwg_handle *h) {
  if (!h) return fail(DWG_ERR_INVALID, "null handle");
  h->chain_open = false;
  CU(cudaSetDevice(h->device));
  CU(cudaMemsetAsync(h->tallies.p, 0, 4 * h->img->ndrow, h->stream));
  h->infer_sweeps = 0;
  h->extra_nsamples.clear();
  return DWG_OK;
}

static int get_tallies_impl(dwg_handle *h, uint64_t *tallies, uint32_t *tallies32, uint64_t *nsamples);
int dwg_get_tallies(dwg_handle *h, uint64_t *tallies, uint64_t *nsamples) {
  return get_tallies_impl(h, tallies, nullptr, nsamples);
}
int dwg_get_tallies32(dwg_handle *h, uint32_t *tallies, uint64_t *nsamples) {
  return get_tallies_impl(h, nullptr, tallies, nsamples);
}
static int get_tallies_impl(dwg_handle *h, uint64_t *tallies, uint32_t *tallies32, uint64_t *nsamples) {
  if (!h) return fail(DWG_ERR_INVALID, "null handle");
  h->chain_open = false;
  const GraphImage &m = *h->img;
  CU(cudaSetDevice(h->device));
  if (h->late_done) CU(cudaStreamWaitEvent(h->stream, h->late_done, 0));  // row_of_value has arrived
  if (!h->row_map_ready && !m.row_of_value.empty()) {  // a deferred replica asked for tallies
    CU(cudaMemcpyAsync(h->row_of_value.p, m.row_of_value.data(), 4 * m.row_of_value.size(), cudaMemcpyHostToDevice, h->stream));
    h->row_map_ready = true;
  }
  const bool timing = getenv("DWG_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    auto n = std::chrono::steady_clock::now();
    if (timing) fprintf(stderr, "[get_tallies] %-24s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t_prev).count());
    t_prev = n;
  };
  // ++agg_nsamples for every sampled variable every sweep (src/gibbs_sampler.h:157,163): implicit on the
  // device, written out here by the host threads
  auto fill_nsamples = [&]() {
    if (!nsamples) return;
    const uint64_t sweeps = h->infer_sweeps;
    const bool all = h->opts.sample_evidence;
    const uint8_t *ev = m.var_is_evid.data();
    const uint64_t *extra = h->extra_nsamples.empty() ? nullptr : h->extra_nsamples.data();
    uint64_t *out = nsamples;
    // (streaming stores measured no faster than plain ones here: 3.4 - 6.2 ms against 3.7 ms for C2's
    // 80 MB on the GPU boxes; the host's store bandwidth is the limit either way)
    parallel_for(m.nvar, [=](uint64_t a, uint64_t b) {
      for (uint64_t v = a; v < b; ++v) out[v] = ((!ev[v] || all) ? sweeps : 0) + (extra ? extra[v] : 0);
    });
    nsamples = nullptr;  // done
  };
  // A page-locked destination gets the counts by DMA from the device (the copy engine writes host
  // memory at the link rate, 52 GB/s measured; sixteen host threads manage 23 GB/s on the same box).
  // The counts depend on the NUMBER of sweeps queued, not on their results: kernel and copy go to the
  // `late` stream and run under sweeps that are still in flight on the handle's stream.
  bool nsamples_dma = false;
  if (nsamples && m.nvar && h->extra_nsamples.empty()) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, nsamples) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
      if (!h->late) CU(cudaStreamCreateWithFlags(&h->late, cudaStreamNonBlocking));
      if (!h->nsamples_dev.p) CU(h->nsamples_dev.alloc(m.nvar));
      if (!h->opts.sample_evidence && !h->evid_abi.p) CU(h->evid_abi.upload(m.var_is_evid, h->late));
      k_fill_nsamples<<<(uint32_t)std::min<uint64_t>((m.nvar + 255) / 256, 8192), 256, 0, h->late>>>(
          h->nsamples_dev.p, h->opts.sample_evidence ? nullptr : h->evid_abi.p, h->infer_sweeps, m.nvar);
      ++g_launches;
      CU(cudaMemcpyAsync(nsamples, h->nsamples_dev.p, 8 * m.nvar, cudaMemcpyDeviceToHost, h->late));
      nsamples = nullptr;  // queued; awaited below
      nsamples_dma = true;
    } else {
      cudaGetLastError();
    }
  }
  if ((tallies || tallies32) && m.nrow) {
    // schedule order -> ABI order on the device, then 4 MB pieces through two small page-locked
    // buffers: the DMA of piece c overlaps the widening of piece c - 1 on the host threads (page-locking
    // a buffer for the whole vector would cost more than the transfer)
    constexpr uint64_t PIECE = 1u << 20;  // tallies per piece
    if (tallies32) {  // a page-locked destination takes the DMA directly
      cudaPointerAttributes attr{};
      if (cudaPointerGetAttributes(&attr, tallies32) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
        if (!h->tallies_abi.p) CU(h->tallies_abi.alloc(m.nrow));
        k_gather_u32<<<(uint32_t)((m.nrow + 255) / 256), 256, 0, h->stream>>>(h->tallies_abi.p, h->tallies.p,
                                                                             h->row_of_value.p, (uint32_t)m.nrow);
        ++g_launches;
        CU(cudaMemcpyAsync(tallies32, h->tallies_abi.p, 4 * m.nrow, cudaMemcpyDeviceToHost, h->stream));
        lap("gather + DMA queued");
        fill_nsamples();  // while the DMA runs
        lap("nsamples (under the DMA)");
        CU(cudaStreamSynchronize(h->stream));
        lap("DMA done");
        tallies32 = nullptr;  // done
      } else {
        cudaGetLastError();
      }
    }
    if (!tallies && !tallies32) goto tallies_done;
    {
    std::lock_guard<std::mutex> pool_lock(g_landing.mu);
    uint32_t *landing = g_landing.get(2 * PIECE);
    if (!landing) return fail(DWG_ERR_CUDA, "cannot page-lock the tally landing buffer");
    if (!h->tallies_abi.p) CU(h->tallies_abi.alloc(m.nrow));
    for (cudaEvent_t &e : h->piece_done)  // (not tied to the buffer: the direct path above allocates it too)
      if (!e) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    lap("buffers");
    k_gather_u32<<<(uint32_t)((m.nrow + 255) / 256), 256, 0, h->stream>>>(h->tallies_abi.p, h->tallies.p,
                                                                         h->row_of_value.p, (uint32_t)m.nrow);
    ++g_launches;
    const uint64_t npiece = (m.nrow + PIECE - 1) / PIECE;
    auto widen = [&](uint64_t c) {
      const uint64_t base = c * PIECE, cnt = std::min(PIECE, m.nrow - base);
      const uint32_t *t = landing + (c & 1) * PIECE;
      if (tallies32) {  // the device's own counter width: a plain copy out of the landing buffer
        uint32_t *out = tallies32 + base;
        parallel_for(cnt, [=](uint64_t a, uint64_t b) { std::memcpy(out + a, t + a, 4 * (b - a)); });
        return;
      }
      uint64_t *out = tallies + base;
      parallel_for(cnt, [=](uint64_t a, uint64_t b) { for (uint64_t i = a; i < b; ++i) out[i] = t[i]; });
    };
    for (uint64_t c = 0; c < npiece; ++c) {
      const uint64_t base = c * PIECE, cnt = std::min(PIECE, m.nrow - base);
      CU(cudaMemcpyAsync(landing + (c & 1) * PIECE, h->tallies_abi.p + base, 4 * cnt, cudaMemcpyDeviceToHost,
                         h->stream));
      CU(cudaEventRecord(h->piece_done[c & 1], h->stream));
      if (c) {
        CU(cudaEventSynchronize(h->piece_done[(c - 1) & 1]));
        widen(c - 1);
      }
    }
    CU(cudaEventSynchronize(h->piece_done[(npiece - 1) & 1]));
    widen(npiece - 1);
    lap("widen");
    }
  }
tallies_done:
  if (nsamples_dma) {
    CU(cudaStreamSynchronize(h->late));
    lap("nsamples (DMA)");
  }
  if (nsamples) {
    fill_nsamples();
    lap("nsamples");
  }
  return DWG_OK;
}

int dwg_add_tallies(dwg_handle *h, const uint64_t *tallies, const uint64_t *nsamples) {
  if (!h || !tallies || !nsamples) return fail(DWG_ERR_INVALID, "null argument");
  h->chain_open = false;
  const GraphImage &m = *h->img;
  CU(cudaSetDevice(h->device));
  // the device's counters are 32 bits wide: a tally is bounded by its variable's sample count, so the
  // counts decide whether the aggregate still fits (refused rather than wrapped)
  uint64_t most = 0;
  for (uint64_t v = 0; v < m.nvar; ++v) most = std::max(most, nsamples[v] + (h->extra_nsamples.empty() ? 0 : h->extra_nsamples[v]));
  if (most + h->infer_sweeps > 0xffffffffull)
    return fail(DWG_ERR_UNSUPPORTED, "dwg_add_tallies: more than 2^32 - 1 samples of one variable (the tally counters are 32 bits wide)");
  std::vector<uint32_t> t(m.ndrow, 0);
  for (uint64_t i = 0; i < m.nrow; ++i) {
    if (tallies[i] > 0xffffffffull) return fail(DWG_ERR_INVALID, "dwg_add_tallies: a tally exceeds its variable's sample count");
    t[m.row_of_value[i]] = (uint32_t)tallies[i];
  }
  DevBuf<uint32_t> tmp;
  CU(tmp.upload(t, h->stream));
  if (m.ndrow) {
    k_add_u32<<<(uint32_t)((m.ndrow + 255) / 256), 256, 0, h->stream>>>(h->tallies.p, tmp.p, (uint32_t)m.ndrow);
    ++g_launches;
  }
  CU(cudaStreamSynchronize(h->stream));
  if (h->extra_nsamples.empty()) h->extra_nsamples.assign(m.nvar, 0);
  for (uint64_t v = 0; v < m.nvar; ++v) h->extra_nsamples[v] += nsamples[v];
  return DWG_OK;
}

int dwg_get_weights(dwg_handle *h, double *out) {
  if (!h || !out) return fail(DWG_ERR_INVALID, "null argument");
  h->chain_open = false;
  const GraphImage &m = *h->img;
  CU(cudaSetDevice(h->device));
  std::vector<double> t(m.nweight);
  CU(cudaMemcpyAsync(t.data(), h->weights.p, 8 * m.nweight, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (uint64_t i = 0; i < m.nweight; ++i) out[m.w_perm[i]] = t[i];
  return DWG_OK;
}

int dwg_set_weights(dwg_handle *h, const double *in) {
  if (!h || !in) return fail(DWG_ERR_INVALID, "null argument");
  h->chain_open = false;
  const GraphImage &m = *h->img;
  CU(cudaSetDevice(h->device));
  std::vector<double> t(m.nweight);
  for (uint64_t i = 0; i < m.nweight; ++i) t[i] = in[m.w_perm[i]];
  CU(cudaMemcpyAsync(h->weights.p, t.data(), 8 * m.nweight, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->prev_weights.p, t.data(), 8 * m.nweight, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DWG_OK;
}

int dwg_set_weight_sum(dwg_handle *h, const double *in) {
  if (!h || !in) return fail(DWG_ERR_INVALID, "null argument");
  h->chain_open = false;
  const GraphImage &m = *h->img;
  CU(cudaSetDevice(h->device));
  std::vector<double> t(m.nweight);
  for (uint64_t i = 0; i < m.nweight; ++i) t[i] = in[m.w_perm[i]];
  CU(cudaMemcpyAsync(h->weights.p, t.data(), 8 * m.nweight, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DWG_OK;
}

int dwg_get_assignments(dwg_handle *h, int evid_chain, uint64_t *out) {
  if (!h || !out) return fail(DWG_ERR_INVALID, "null argument");
  h->chain_open = false;
  const GraphImage &m = *h->img;
  CU(cudaSetDevice(h->device));
  const void *src = evid_chain ? h->ds.assign_evid : h->ds.assign_free;
  if (m.assign_bytes == 1) {
    std::vector<uint8_t> t(m.nvar_pad);
    CU(cudaMemcpyAsync(t.data(), src, m.nvar_pad, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (uint64_t v = 0; v < m.nvar; ++v) out[v] = t[m.inv[v]];
  } else {
    std::vector<uint32_t> t(m.nvar_pad);
    CU(cudaMemcpyAsync(t.data(), src, 4 * m.nvar_pad, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (uint64_t v = 0; v < m.nvar; ++v) out[v] = t[m.inv[v]];
  }
  return DWG_OK;
}

int dwg_set_assignments(dwg_handle *h, int evid_chain, const uint64_t *in) {
  if (!h || !in) return fail(DWG_ERR_INVALID, "null argument");
  h->chain_open = false;
  const GraphImage &m = *h->img;
  CU(cudaSetDevice(h->device));
  void *dst = evid_chain ? h->ds.assign_evid : h->ds.assign_free;
  if (m.assign_bytes == 1) {
    std::vector<uint8_t> t(m.nvar_pad, 0);
    for (uint64_t v = 0; v < m.nvar; ++v) t[m.inv[v]] = (uint8_t)in[v];
    CU(cudaMemcpyAsync(dst, t.data(), m.nvar_pad, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  } else {
    std::vector<uint32_t> t(m.nvar_pad, 0);
    for (uint64_t v = 0; v < m.nvar; ++v) t[m.inv[v]] = (uint32_t)in[v];
    CU(cudaMemcpyAsync(dst, t.data(), 4 * m.nvar_pad, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return DWG_OK;
}

int dwg_finish_weight_average(dwg_handle *h, int n_replicas, double *lmax, double *l2) {
  if (!h || n_replicas < 1) return fail(DWG_ERR_INVALID, "bad argument");
  h->chain_open = false;
  const uint32_t n = (uint32_t)h->img->nweight;
  CU(cudaSetDevice(h->device));
  CU(cudaMemsetAsync(h->norms.p, 0, 16, h->stream));
  if (n) {
    uint32_t blocks = std::min<uint32_t>((n + 255) / 256, (uint32_t)h->num_sms * 4);
    k_finish_average<<<blocks, 256, 0, h->stream>>>(h->weights.p, h->prev_weights.p, n,
                                                    1.0 / n_replicas, n_replicas > 1, h->norms.p);
    ++g_launches;
    CU(cudaGetLastError());
  }
  if (lmax || l2) {
    double host[2];
    CU(cudaMemcpyAsync(host, h->norms.p, 16, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (lmax) *lmax = host[0];
    if (l2) *l2 = host[1];  // sum of squares; the caller prints sqrt(l2)/stepsize
  }
  return DWG_OK;
}

// ---- NCCL, dlopen'ed so that libdwg.so itself has no NCCL dependency ----------------
namespace {
struct NcclApi {
  int (*CommInitAll)(void **, int, const int *) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Reduce)(const void *, void *, size_t, int, int, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::string err;
  bool ok = false;
};
constexpr int NCCL_UINT32 = 3, NCCL_FLOAT64 = 8, NCCL_SUM = 0;  // nccl.h enums

NcclApi &nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void *lib = nullptr;
  if (!dlsym(RTLD_DEFAULT, "ncclAllReduce")) {  // not already in the process (e.g. via torch)
    const char *env = getenv("DWG_NCCL_LIB");
    if (env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
      api.err = "libnccl.so.2 not loadable (set DWG_NCCL_LIB or LD_LIBRARY_PATH)";
      return api;
    }
  }
  void *h = lib ? lib : RTLD_DEFAULT;
#define SYM(name) api.name = (decltype(api.name))dlsym(h, "nccl" #name)
  SYM(CommInitAll); SYM(CommDestroy); SYM(AllReduce); SYM(Reduce); SYM(GroupStart); SYM(GroupEnd);
  SYM(GetErrorString);
#undef SYM
  api.ok = api.CommInitAll && api.CommDestroy && api.AllReduce && api.Reduce && api.GroupStart && api.GroupEnd;
  if (!api.ok) api.err = "libnccl is missing expected symbols";
  return api;
}
int nccl_fail(int rc, const char *what) {
  NcclApi &n = nccl();
  return fail(DWG_ERR_NCCL, std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "nccl error"));
}
}  // namespace

int dwg_comm_init_all(int n, const int *devices, void **comms_out) {
  if (n < 1 || !devices || !comms_out) return fail(DWG_ERR_INVALID, "bad argument");
  NcclApi &api = nccl();
  if (!api.ok) return fail(DWG_ERR_NCCL, api.err);
  int rc = api.CommInitAll(comms_out, n, devices);
  if (rc) return nccl_fail(rc, "ncclCommInitAll");
  return DWG_OK;
}

int dwg_comm_destroy(void *comm) {
  if (!comm) return DWG_OK;
  NcclApi &api = nccl();
  if (!api.ok) return fail(DWG_ERR_NCCL, api.err);
  api.CommDestroy(comm);
  return DWG_OK;
}

int dwg_average_weights_nccl(dwg_handle *h, void *comm, int n_replicas, double *lmax, double *l2) {
  if (!h) return fail(DWG_ERR_INVALID, "null handle");
  h->chain_open = false;
  if (n_replicas > 1) {
    if (!comm) return fail(DWG_ERR_INVALID, "null communicator");
    NcclApi &api = nccl();
    if (!api.ok) return fail(DWG_ERR_NCCL, api.err);
    CU(cudaSetDevice(h->device));
    int rc = api.AllReduce(h->weights.p, h->weights.p, h->img->nweight, NCCL_FLOAT64, NCCL_SUM, comm, h->stream);
    if (rc) return nccl_fail(rc, "ncclAllReduce");
  }
  return dwg_finish_weight_average(h, n_replicas, lmax, l2);
}

int dwg_average_weights_multi(dwg_handle *const *hs, void *const *comms, int n, double *lmax, double *l2) {
  if (!hs || n < 1) return fail(DWG_ERR_INVALID, "bad argument");
  const size_t nw = hs[0]->img->nweight;
  if (n > 1 && comms) {
    NcclApi &api = nccl();
    if (!api.ok) return fail(DWG_ERR_NCCL, api.err);
    int rc = api.GroupStart();
    if (rc) return nccl_fail(rc, "ncclGroupStart");
    for (int i = 0; i < n; ++i) {
      rc = api.AllReduce(hs[i]->weights.p, hs[i]->weights.p, nw, NCCL_FLOAT64, NCCL_SUM, comms[i], hs[i]->stream);
      if (rc) { api.GroupEnd(); return nccl_fail(rc, "ncclAllReduce"); }
    }
    rc = api.GroupEnd();
    if (rc) return nccl_fail(rc, "ncclGroupEnd");
  } else if (n > 1) {
    // no NCCL: sum on the host (what the reference's main thread does, dimmwitted.cc:214-215)
    std::vector<double> sum(nw, 0.0), tmp(nw);
    for (int i = 0; i < n; ++i) {
      int rc = dwg_get_weights(hs[i], tmp.data());
      if (rc) return rc;
      for (size_t j = 0; j < nw; ++j) sum[j] += tmp[j];
    }
    // dwg_get_weights hands out ABI order; the device vector is in internal (hottest first) order:
    // the sum goes back through dwg_set_weight_sum, which applies the same permutation
    for (int i = 0; i < n; ++i) {
      int rc = dwg_set_weight_sum(hs[i], sum.data());
      if (rc) return rc;
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    int rc = dwg_finish_weight_average(hs[i], n, i == 0 ? lmax : nullptr, i == 0 ? l2 : nullptr);
    if (rc) return rc;
  }
  return DWG_OK;
}

int dwg_reduce_tallies_multi(dwg_handle *const *hs, void *const *comms, int n) {
  if (!hs || n < 1) return fail(DWG_ERR_INVALID, "bad argument");
  const GraphImage &m = *hs[0]->img;
  if (n > 1 && comms) {
    NcclApi &api = nccl();
    if (!api.ok) return fail(DWG_ERR_NCCL, api.err);
    int rc = api.GroupStart();
    if (rc) return nccl_fail(rc, "ncclGroupStart");
    for (int i = 0; i < n; ++i) {
      rc = api.Reduce(hs[i]->tallies.p, hs[i]->tallies.p, m.ndrow, NCCL_UINT32, NCCL_SUM, 0, comms[i], hs[i]->stream);
      if (rc) { api.GroupEnd(); return nccl_fail(rc, "ncclReduce"); }
    }
    rc = api.GroupEnd();
    if (rc) return nccl_fail(rc, "ncclGroupEnd");
    for (int i = 1; i < n; ++i) hs[0]->infer_sweeps += hs[i]->infer_sweeps;
  } else if (n > 1) {
    std::vector<uint64_t> t(m.nrow), ns(m.nvar);
    for (int i = 1; i < n; ++i) {
      int rc = dwg_get_tallies(hs[i], t.data(), ns.data());
      if (rc) return rc;
      rc = dwg_add_tallies(hs[0], t.data(), ns.data());
      if (rc) return rc;
    }
  }
  return DWG_OK;
}

// ---- parity hooks --------------------------------------------------------------------
int dwg_eval_potentials(dwg_handle *h, uint64_t n, const uint64_t *vids, const uint64_t *proposals,
                        int evid_chain, double *out) {
  if (!h || !vids || !proposals || !out) return fail(DWG_ERR_INVALID, "null argument");
  h->chain_open = false;
  const GraphImage &m = *h->img;
  CU(cudaSetDevice(h->device));
  {
    int rc = ensure_csr(h);
    if (rc) return rc;
  }
  std::vector<uint32_t> ds(n), ps(n);
  for (uint64_t i = 0; i < n; ++i) {
    if (vids[i] >= m.nvar) return fail(DWG_ERR_INVALID, "variable id out of range");
    uint64_t k = proposals[i];
    if (m.var_dtype[vids[i]] && k >= m.value_base[vids[i] + 1] - m.value_base[vids[i]])
      return fail(DWG_ERR_INVALID, "proposal out of range");
    ds[i] = m.inv[vids[i]];
    ps[i] = (uint32_t)std::min<uint64_t>(k, 0xffffffffu);
  }
  DevBuf<uint32_t> dd, dp;
  DevBuf<double> dout;
  CU(dd.upload(ds, h->stream));
  CU(dp.upload(ps, h->stream));
  CU(dout.alloc(n));
  if (n) {
    uint32_t blocks = (uint32_t)((n + 127) / 128);
    if (m.assign_bytes == 1)
      k_eval_potentials<uint8_t><<<blocks, 128, 0, h->stream>>>(h->dg, h->ds, (uint32_t)n, dd.p, dp.p, evid_chain, dout.p);
    else
      k_eval_potentials<uint32_t><<<blocks, 128, 0, h->stream>>>(h->dg, h->ds, (uint32_t)n, dd.p, dp.p, evid_chain, dout.p);
    ++g_launches;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout.p, 8 * n, cudaMemcpyDeviceToHost, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  return DWG_OK;
}

int dwg_draw(int device, uint64_t n, const double *pots, uint64_t ld, const uint32_t *card,
             const double *r, uint32_t *out) {
  if (!pots || !card || !r || !out) return fail(DWG_ERR_INVALID, "null argument");
  CU(cudaSetDevice(device));
  DevBuf<double> dp, dr;
  DevBuf<uint32_t> dc, dout;
  CU(dp.alloc(n * ld)); CU(dr.alloc(n)); CU(dc.alloc(n)); CU(dout.alloc(n));
  CU(cudaMemcpy(dp.p, pots, 8 * n * ld, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dr.p, r, 8 * n, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dc.p, card, 4 * n, cudaMemcpyHostToDevice));
  if (n) {
    k_draw<<<(uint32_t)((n + 127) / 128), 128>>>((uint32_t)n, dp.p, (uint32_t)ld, dc.p, dr.p, dout.p);
    ++g_launches;
    CU(cudaGetLastError());
  }
  CU(cudaMemcpy(out, dout.p, 4 * n, cudaMemcpyDeviceToHost));
  return DWG_OK;
}

int dwg_logadd(int device, uint64_t n, const double *a, const double *b, double *out) {
  if (!a || !b || !out) return fail(DWG_ERR_INVALID, "null argument");
  CU(cudaSetDevice(device));
  DevBuf<double> da, db, dout;
  CU(da.alloc(n)); CU(db.alloc(n)); CU(dout.alloc(n));
  CU(cudaMemcpy(da.p, a, 8 * n, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(db.p, b, 8 * n, cudaMemcpyHostToDevice));
  if (n) {
    k_logadd<<<(uint32_t)((n + 127) / 128), 128>>>((uint32_t)n, da.p, db.p, dout.p);
    ++g_launches;
    CU(cudaGetLastError());
  }
  CU(cudaMemcpy(out, dout.p, 8 * n, cudaMemcpyDeviceToHost));
  return DWG_OK;
}

int dwg_debug_set_uniforms(dwg_handle *h, const double *r_free, const double *r_evid) {
  if (!h) return fail(DWG_ERR_INVALID, "null handle");
  h->chain_open = false;
  const GraphImage &m = *h->img;
  CU(cudaSetDevice(h->device));
  if (!r_free) { h->explicit_r = false; return DWG_OK; }
  std::vector<double> a(m.nvar_pad, 0.0), b(m.nvar_pad, 0.0);
  for (uint64_t v = 0; v < m.nvar; ++v) {
    a[m.inv[v]] = r_free[v];
    if (r_evid) b[m.inv[v]] = r_evid[v];
  }
  if (!h->r_free.p) { CU(h->r_free.alloc(m.nvar_pad)); CU(h->r_evid.alloc(m.nvar_pad)); }
  CU(cudaMemcpyAsync(h->r_free.p, a.data(), 8 * m.nvar_pad, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->r_evid.p, b.data(), 8 * m.nvar_pad, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->explicit_r = true;
  return DWG_OK;
}

int dwg_debug_set_sequential(dwg_handle *h, int on) {
  if (!h) return fail(DWG_ERR_INVALID, "null handle");
  h->chain_open = false;
  h->sequential = on != 0;
  return DWG_OK;
}

}  // extern "C"
