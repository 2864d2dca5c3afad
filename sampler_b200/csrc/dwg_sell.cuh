// dwg_sell.cuh -- the production sweep kernels for Boolean graphs (sm_100a): one warp per
// 32-variable slice of the sliced-ELL blob (graph_image.h), lane = variable.
//
//   k_sell_infer_staged  GibbsSamplerThread::sample   (src/gibbs_sampler.cc:65-69,
//   k_sell_infer                                        src/gibbs_sampler.h:151-169,192-215)
//   k_sell_learn_live    GibbsSamplerThread::sample_sgd (src/gibbs_sampler.cc:71-75,
//                          src/gibbs_sampler.h:127-149,171-190; sgd_on_variable/-factor,
//                          src/factor_graph.cc:243-314; update_weight, src/inference_result.h:66-85)
//   k_apply_weights      the batch weight update of the CSR (categorical) learning launches
//
// Memory behaviour (HBM-bound by design): a slice is one contiguous run of words, read exactly
// once per sweep.  The *_staged kernels (the default) move it with one 1-D bulk-async copy
// (cp.async.bulk + mbarrier, the TMA unit) into a per-warp ring in shared memory, one or more
// slices ahead of the one being sampled, and read records with conflict-free LDS; the direct
// kernels (fallback for slices longer than a ring stage) read it with warp-wide 128 B global
// loads issued before their first dependent use.  Per variable besides the records: 1 B chain
// write, a tally RED; evidence / padding / evidence-value bits ride in the slice header.
// Gathers that stay on chip: weights (shared memory: a signed table -w | 0 | +w in the staged
// inference kernel, a plain one elsewhere; ids beyond it through L1/L2) and the u8 chain
// values of other variables (L2).  The fast paths are branch-free: a unary record carries its
// two sign codes, a record of a two-variable factor carries the factor's whole 2x2 sign table,
// so no factor-function switch runs per edge.  Persistent CTAs (2 per SM) walk the slices of a
// launch with a grid stride, so the weight table is staged once per CTA; consecutive launches
// are chained by programmatic dependent launch (their prologues overlap the previous tail).
// No tensor cores: O(1) flop per 4..12 B of edge data.
#ifndef DWG_SELL_CUH_
#define DWG_SELL_CUH_
#include "dwg_device.cuh"

namespace dwg {

constexpr int SELL_THREADS = 512;  // 16 warps = 16 slices in flight per CTA
constexpr int SELL_MCHUNK = 4;     // multi records / others words loaded together

struct SellGraph {
  uint32_t slice_begin, slice_end;  // slices of this launch (one colour class)
  uint32_t slice_stride;            // 1, or the number of learning chunks (begin carries the phase)
  const uint32_t *sl_off;           // [nslice+1] first word of a slice / 8
  const uint32_t *blob;
  uint32_t nweight, n_hot;
  uint32_t n_tab;  // staged kernel: weights in its signed table (<= SELL_MAX_TAB)
};

struct SellState {
  uint8_t *assign_free, *assign_evid;
  const double *weights;  // frozen for the duration of a launch
  double *grad_acc;       // [nweight] sum of gradients of this launch (learning)
  uint32_t *tallies;
  // hubs cut into parts (SF_PART, graph_image.h), per slice; null when the graph has none
  double *part_sums;      // [4 * nslice] lane-combined potentials of a part: free pos / neg, evid pos / neg
  uint32_t *part_done;    // [nslice] at a hub's first slice: parts finished in this sweep
};

struct SliceHeader {
  uint32_t n_un, n_mu, n_oc, evid, pad, evval, flags;
  const uint32_t *un, *mu, *oth;  // record sections (feature values follow each)
};

template <int FV>
__device__ __forceinline__ SliceHeader sell_header(const SellGraph &g, uint32_t s) {
  constexpr uint32_t FVW = FV == FV_F32 ? 1 : FV == FV_F64 ? 2 : 0;
  const uint32_t *b = g.blob + (size_t)__ldg(g.sl_off + s) * 8;
  const uint4 h0 = __ldg(reinterpret_cast<const uint4 *>(b));
  const uint4 h1 = __ldg(reinterpret_cast<const uint4 *>(b) + 1);
  SliceHeader h;
  h.n_un = h0.x; h.n_mu = h0.y; h.n_oc = h0.z; h.evid = h0.w;
  h.pad = h1.x; h.evval = h1.y; h.flags = h1.z;
  h.un = b + SH_WORDS;
  h.mu = h.un + (size_t)32 * h.n_un * (1 + FVW);
  h.oth = h.mu + (size_t)32 * h.n_mu * (1 + FVW);
  return h;
}

// how a slice is read: cached loads for learning (its SGD pass walks the slice a second time
// and finds it in L1), streaming (evict-first) loads for the direct inference sweep, plain
// loads when the slice was staged in shared memory by a bulk-async copy
// LD_CACHED and LD_SMEM_LIVE mark a LEARNING sweep (its weights and chains are live, see sell_live)
constexpr int LD_CACHED = 0, LD_STREAM = 1, LD_SMEM = 2, LD_SMEM_LIVE = 3;
template <int LD, typename T>
__device__ __forceinline__ T sell_ldt(const T *p) {
  return (LD == LD_SMEM || LD == LD_SMEM_LIVE) ? *p : LD == LD_STREAM ? __ldcs(p) : __ldg(p);
}
__host__ __device__ constexpr bool sell_live(int LD) { return LD == LD_CACHED || LD == LD_SMEM_LIVE; }
template <int LD>
__device__ __forceinline__ uint32_t sell_ld(const uint32_t *p) { return sell_ldt<LD, uint32_t>(p); }

// feature value of record (column j, lane) of a section whose records start at `sec` and
// hold `ncol` columns
template <int FV, int LD>
__device__ __forceinline__ double sell_fv(const uint32_t *sec, uint32_t ncol, uint32_t j, uint32_t lane) {
  if (FV == FV_F32) {
    const float *p = reinterpret_cast<const float *>(sec + (size_t)32 * ncol) + (size_t)32 * j + lane;
    return (double)sell_ldt<LD, float>(p);
  }
  if (FV == FV_F64) {
    const double *p = reinterpret_cast<const double *>(sec + (size_t)32 * ncol) + (size_t)32 * j + lane;
    return sell_ldt<LD, double>(p);
  }
  return 1.0;
}

// chain value of another variable: a learning sweep is Hogwild (the variable may be resampled by
// another SM during the launch), so it reads L2; a colour class of an inference sweep never
// writes what it gathers
template <int LD>
__device__ __forceinline__ uint32_t sell_chain(const uint8_t *a, uint32_t d) {
  return sell_live(LD) ? (uint32_t)__ldcg(a + d) : (uint32_t)a[d];
}

// SF_WIDE slices: the 32 lanes hold shares of ONE variable's records; their partial potentials are
// combined by a butterfly of shuffles (a fixed order: the same sum on every lane, run to run)
__device__ __forceinline__ double sell_wide_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = __dadd_rn(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

// A part of a hub is done: store its (lane-combined) sums and count it; the warp that counts LAST
// adds the parts' sums in a fixed order (lane q takes parts q, q + 32, ..., then the butterfly) --
// the same result whichever warp that is.  nv: values per part (2 inference, 4 learning).
// Returns true on the finishing warp, with v[] = the hub's sums.
template <int NV>
__device__ __forceinline__ bool sell_part_finish(const SellState &st, uint32_t s, uint32_t flags, uint32_t lane, double (&v)[NV]) {
  const uint32_t q = (flags >> 4) & 0x3fffu, n = flags >> 18, sf = s - q;
  uint32_t old = 0;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) __stcg(st.part_sums + 4 * (size_t)s + i, v[i]);
    __threadfence();
    old = atomicAdd(st.part_done + sf, 1u);
  }
  old = __shfl_sync(0xffffffffu, old, 0);
  if (old != n - 1) return false;
  __threadfence();
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  for (uint32_t k = lane; k < n; k += 32)
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = __dadd_rn(acc[i], __ldcg(st.part_sums + 4 * (size_t)(sf + k) + i));
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = sell_wide_sum(acc[i]);
  if (lane == 0) st.part_done[sf] = 0;  // ready for the next sweep
  return true;
}

// sign code (0,1,2) -> the term  sign * t  with t = weight * feature value
__device__ __forceinline__ double sell_term(uint32_t code, double t) {
  return code == 2u ? t : (code == 0u ? -t : 0.0);
}

// the weights of a learning sweep beyond the shared-memory copy are LIVE (other CTAs move them
// with atomics while the launch runs), so they are read from L2, not through the non-coherent path
template <int LD>
__device__ __forceinline__ double sell_weight(const SellGraph &g, const double *w_s, const double *w, uint32_t wid) {
  return wid < g.n_hot ? w_s[wid] : (sell_live(LD) ? __ldcg(w + wid) : __ldg(w + wid));
}

__device__ __forceinline__ void sell_stage_weights(const SellGraph &g, const double *w, double *w_s) {
  for (uint32_t i = threadIdx.x; i < g.n_hot; i += blockDim.x) w_s[i] = __ldg(w + i);
}

// Boolean draw (src/gibbs_sampler.h:198-215) with an fp32 screen in front of the fp64 test
//     r * (1 + exp(neg - pos)) < 1
// The fp32 value is within 1e-5 relative of the fp64 one for |neg - pos| <= 64, so whenever it
// is further than 1e-3 from 1 both agree; otherwise (about one draw in a thousand) the exact
// fp64 expression decides.  The result is the reference's bit for bit.
__device__ __forceinline__ uint32_t sell_draw_boolean(double pos, double neg, double r) {
  const double d = __dsub_rn(neg, pos);
  const float df = (float)d, rf = (float)r;
  const float x = fmaf(rf, __expf(df), rf);
  if (fabsf(x - 1.0f) > 1e-3f && fabsf(df) <= 64.0f) return x < 1.0f ? 1u : 0u;
  return __dmul_rn(r, __dadd_rn(1.0, exp(d))) < 1.0 ? 1u : 0u;
}

// ---- unary records: pos/neg += weight * (sign * feature_value), src/factor_graph.h:140-142 ----
template <int FV, int LD, bool ANTISYM>
__device__ __forceinline__ void sell_unary(const SellGraph &g, const SliceHeader &h, const double *w_s,
                                           const double *w, uint32_t lane, double &pos, double &neg) {
  // (a learning sweep reading a long slice from L2 / HBM -- LD_CACHED, 128-register kernel -- keeps more loads in flight)
  constexpr int U = FV == FV_F64 ? (LD == LD_CACHED ? 10 : 4) : 8;
  for (uint32_t j0 = 0; j0 < h.n_un; j0 += U) {
    uint32_t rec[U];
    double fv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = j0 + u < h.n_un;
      rec[u] = ok ? sell_ld<LD>(h.un + (size_t)32 * (j0 + u) + lane) : SELL_UNARY_NOP;
      fv[u] = (FV != FV_ONE && ok) ? sell_fv<FV, LD>(h.un, h.n_un, j0 + u, lane) : 1.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double wt = sell_weight<LD>(g, w_s, w, sell_unary_wid(rec[u]));
      const double t = FV == FV_ONE ? wt : __dmul_rn(wt, fv[u]);  // w * (+-1 * fv) == +-(w * fv) exactly
      pos = __dadd_rn(pos, sell_term(rec[u] & 3u, t));
      if (!ANTISYM) neg = __dadd_rn(neg, sell_term(rec[u] >> 26, t));
    }
  }
  if (ANTISYM) neg = -pos;  // round-to-nearest is sign-symmetric: the mirrored sum, bit for bit
}

// ---- multi records of a general slice: CSR tag semantics, per-lane cursor into the others ----
// Signs of one record when the variable is 0 / 1, the other variables read from chain `a`.
template <int LD>
__device__ __forceinline__ void sell_general_signs(const SliceHeader &h, uint32_t tag, uint32_t self,
                                                   uint32_t lane, uint32_t &oc, const uint8_t *a,
                                                   double &s0, double &s1) {
  const uint32_t fc = tag & TAG_FUNC_MASK, form = tag >> TAG_FORM_SHIFT;
  const uint32_t *o = h.oth + lane;
  if (form != FORM_GENERAL) {
    const uint32_t self_eq = (tag & TAG_SELF_EQ1) ? 1u : 0u;
    const bool self_head = tag & TAG_SELF_HEAD;
    uint32_t nsb = 0;
    bool head_oth = false;
    for (uint32_t j = 0; j < form; ++j) {
      const uint32_t w = sell_ld<LD>(o + (size_t)(oc + j) * 32);
      const bool s = sell_chain<LD>(a, w >> 1) == (w & 1u);
      if (!self_head && j == form - 1) head_oth = s;
      else nsb += s;
    }
    oc += form;
    if (self_head) {
      s0 = dev_sign(fc, form + 1, nsb, self_eq == 0u);
      s1 = dev_sign(fc, form + 1, nsb, self_eq == 1u);
    } else {
      s0 = dev_sign(fc, form + 1, nsb + (self_eq == 0u), head_oth);
      s1 = dev_sign(fc, form + 1, nsb + (self_eq == 1u), head_oth);
    }
    return;
  }
  const uint32_t n = sell_ld<LD>(o + (size_t)oc * 32);
  uint32_t nsb0 = 0, nsb1 = 0;
  bool h0 = false, h1 = false;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t w = sell_ld<LD>(o + (size_t)(oc + 1 + i) * 32);
    const uint32_t d = w >> 1, eq = w & 1u;
    bool t0, t1;
    if (d == self) { t0 = eq == 0u; t1 = eq == 1u; }
    else { t0 = t1 = sell_chain<LD>(a, d) == eq; }
    if (i == n - 1) { h0 = t0; h1 = t1; }
    else { nsb0 += t0; nsb1 += t1; }
  }
  oc += 1 + n;
  s0 = dev_sign(fc, n, nsb0, h0);
  s1 = dev_sign(fc, n, nsb1, h1);
}

// number of others words behind a general-slice record (to skip it)
template <int LD>
__device__ __forceinline__ uint32_t sell_general_span(const SliceHeader &h, uint32_t tag, uint32_t lane, uint32_t oc) {
  const uint32_t form = tag >> TAG_FORM_SHIFT;
  return form != FORM_GENERAL ? form : 1u + sell_ld<LD>(h.oth + (size_t)oc * 32 + lane);
}

template <int FV, int LD>
__device__ __forceinline__ void sell_general_multi(const SellGraph &g, const SliceHeader &h, const double *w_s,
                                                   const double *w, uint32_t lane, uint32_t self,
                                                   const uint8_t *a, double &pos, double &neg) {
  uint32_t oc = 0;
  for (uint32_t j = 0; j < h.n_mu; ++j) {
    const uint32_t rec = sell_ld<LD>(h.mu + (size_t)32 * j + lane);
    const uint32_t tag = rec >> 24;
    if ((tag >> TAG_FORM_SHIFT) == 0) continue;  // padding
    const double wt = sell_weight<LD>(g, w_s, w, rec & SELL_WID_MASK);
    const double fv = sell_fv<FV, LD>(h.mu, h.n_mu, j, lane);
    double s0, s1;
    sell_general_signs<LD>(h, tag, self, lane, oc, a, s0, s1);
    pos = dev_acc(pos, wt, s1, fv);
    neg = dev_acc(neg, wt, s0, fv);
  }
}

// ---- multi records of a binary slice: record j owns others word j; the tag is the factor's
// sign table over (value of this variable, other variable satisfied) ----
struct BinaryChunk {
  uint32_t rec[SELL_MCHUNK], ow[SELL_MCHUNK];
};

template <int LD>
__device__ __forceinline__ void sell_binary_load(const SliceHeader &h, uint32_t j0, uint32_t lane, BinaryChunk &c) {
#pragma unroll
  for (int u = 0; u < SELL_MCHUNK; ++u) {
    const bool ok = j0 + u < h.n_mu;
    c.rec[u] = ok ? sell_ld<LD>(h.mu + (size_t)32 * (j0 + u) + lane) : (SELL_BINARY_NOP << 24);
    c.ow[u] = ok ? sell_ld<LD>(h.oth + (size_t)32 * (j0 + u) + lane) : 0u;
  }
}

// bit u: the other variable of record u satisfies its equal_to on chain `a`
template <int LD>
__device__ __forceinline__ uint32_t sell_binary_sat(const BinaryChunk &c, const uint8_t *a) {
  uint32_t av[SELL_MCHUNK], sat = 0;
#pragma unroll
  for (int u = 0; u < SELL_MCHUNK; ++u) av[u] = sell_chain<LD>(a, c.ow[u] >> 1);
#pragma unroll
  for (int u = 0; u < SELL_MCHUNK; ++u) sat |= (av[u] == (c.ow[u] & 1u) ? 1u : 0u) << u;
  return sat;
}

template <int FV, int LD>
__device__ __forceinline__ void sell_binary_acc(const SellGraph &g, const SliceHeader &h, const double *w_s,
                                                const double *w, uint32_t j0, uint32_t lane,
                                                const BinaryChunk &c, uint32_t sat, double &pos, double &neg) {
#pragma unroll
  for (int u = 0; u < SELL_MCHUNK; ++u) {
    const double wt = sell_weight<LD>(g, w_s, w, c.rec[u] & SELL_WID_MASK);
    double t = wt;
    if (FV != FV_ONE) t = __dmul_rn(wt, j0 + u < h.n_mu ? sell_fv<FV, LD>(h.mu, h.n_mu, j0 + u, lane) : 0.0);
    const uint32_t tab = (c.rec[u] >> 24) >> (2u * ((sat >> u) & 1u));
    pos = __dadd_rn(pos, sell_term((tab >> 4) & 3u, t));
    neg = __dadd_rn(neg, sell_term(tab & 3u, t));
  }
}

// The multi-variable records of a slice added to (pos, neg), the other variables read from chain
// `a`.  first / first_sat: the slice's first binary chunk when the caller already loaded it.
template <int FV, int LD>
__device__ __forceinline__ void sell_multi(const SellGraph &g, const SliceHeader &h, const double *w_s,
                                           const double *w, uint32_t lane, uint32_t self, const uint8_t *a,
                                           const BinaryChunk *first, uint32_t first_sat, double &pos, double &neg) {
  if (h.flags & SF_BINARY) {
    BinaryChunk c;
    uint32_t sat;
    if (first) { c = *first; sat = first_sat; }
    else { sell_binary_load<LD>(h, 0, lane, c); sat = sell_binary_sat<LD>(c, a); }
    sell_binary_acc<FV, LD>(g, h, w_s, w, 0, lane, c, sat, pos, neg);
    for (uint32_t j0 = SELL_MCHUNK; j0 < h.n_mu; j0 += SELL_MCHUNK) {
      sell_binary_load<LD>(h, j0, lane, c);
      sat = sell_binary_sat<LD>(c, a);
      sell_binary_acc<FV, LD>(g, h, w_s, w, j0, lane, c, sat, pos, neg);
    }
  } else if (h.n_mu) {
    sell_general_multi<FV, LD>(g, h, w_s, w, lane, self, a, pos, neg);
  }
}

// FactorGraph::potential for proposals 1 and 0 of the lane's variable on chain `a`
// (src/factor_graph.h:127-145).  The slice's first binary chunk is loaded ahead of the unary
// records so that its gathers overlap them.  upos / uneg (optional): the sums after the unary
// records -- they do not depend on the chain, so a second chain continues from them (sell_multi).
template <bool MULTI, int FV, int LD>
__device__ __forceinline__ void sell_potentials(const SellGraph &g, const SliceHeader &h, const double *w_s,
                                                const double *w, uint32_t lane, uint32_t self,
                                                const uint8_t *a, double &pos, double &neg,
                                                double *upos = nullptr, double *uneg = nullptr) {
  pos = 0.0; neg = 0.0;
  BinaryChunk c;
  uint32_t sat = 0;
  const bool binary = MULTI && (h.flags & SF_BINARY);
  if (binary) {
    sell_binary_load<LD>(h, 0, lane, c);
    sat = sell_binary_sat<LD>(c, a);
  }
  if (h.flags & SF_ANTISYM) sell_unary<FV, LD, true>(g, h, w_s, w, lane, pos, neg);
  else sell_unary<FV, LD, false>(g, h, w_s, w, lane, pos, neg);
  if (upos) { *upos = pos; *uneg = neg; }
  if (!MULTI) return;
  sell_multi<FV, LD>(g, h, w_s, w, lane, self, a, binary ? &c : nullptr, sat, pos, neg);
}

// The draw of an SF_WIDE slice (inference): combine the lanes' shares, for a part of a hub store
// them and let the warp that finishes the hub's last part go on, then lane 0 draws for the variable.
// Only kernels instantiated with WIDE = true contain it: graphs without wide slices run the very
// code they ran before wide slices existed.
__device__ __forceinline__ void sell_wide_draw(const SellState &st, const SweepParams &p, uint32_t s, uint32_t flags,
                                               uint32_t lane, uint32_t active, double pos, double neg) {
  pos = sell_wide_sum(pos); neg = sell_wide_sum(neg);
  uint32_t sfirst = s;
  if (flags & SF_PART) {
    double v[2] = {pos, neg};
    if (!sell_part_finish<2>(st, s, flags, lane, v)) return;
    pos = v[0]; neg = v[1];
    sfirst = s - ((flags >> 4) & 0x3fffu);
  }
  if (lane != 0 || !(active & 1u)) return;
  const uint32_t d = sfirst * 32u;
  double r0, r1;
  dev_uniforms(p, d, r0, r1);
  const uint32_t prop = sell_draw_boolean(pos, neg, r0);
  st.assign_evid[d] = (uint8_t)prop;
  if (prop) atomicAdd(st.tallies + d, 1u);
}
// slot of the variable a wide slice belongs to (general records compare positions with it)
template <bool WIDE>
__device__ __forceinline__ uint32_t sell_self_slot(uint32_t s, uint32_t flags, uint32_t lane) {
  if (!WIDE || !(flags & SF_WIDE)) return s * 32u + lane;
  return (s - ((flags & SF_PART) ? (flags >> 4) & 0x3fffu : 0u)) * 32u;
}

template <bool MULTI, int FV, bool WIDE>
__global__ void __launch_bounds__(SELL_THREADS, 2)
k_sell_infer(const SellGraph g, const SellState st, const SweepParams p) {
  extern __shared__ double w_s[];
  sell_stage_weights(g, st.weights, w_s);
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint64_t s64 = g.slice_begin + (uint64_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * g.slice_stride;
       s64 < g.slice_end; s64 += (uint64_t)warps * g.slice_stride) {
    const uint32_t s = (uint32_t)s64;
    const SliceHeader h = sell_header<FV>(g, s);
    // gibbs_sampler.h:157 -- slices are sorted by evidence: one with nothing to sample is skipped
    const uint32_t active = ~h.pad & (p.sample_evidence ? 0xffffffffu : ~h.evid);
    if (!active) continue;
    const uint32_t d = s * 32u + lane;
    double pos, neg;
    sell_potentials<MULTI, FV, LD_STREAM>(g, h, w_s, st.weights, lane, sell_self_slot<WIDE>(s, h.flags, lane), st.assign_evid, pos, neg);
    if (WIDE && (h.flags & SF_WIDE)) { sell_wide_draw(st, p, s, h.flags, lane, active, pos, neg); continue; }  // warp per variable
    if (!((active >> lane) & 1u)) continue;
    double r0, r1;
    dev_uniforms(p, d, r0, r1);
    const uint32_t prop = sell_draw_boolean(pos, neg, r0);
    st.assign_evid[d] = (uint8_t)prop;
    if (prop) atomicAdd(st.tallies + d, 1u);  // gibbs_sampler.h:163-167 (agg_nsamples is implicit); a fire-and-forget RED, the warp does not wait for the tally
  }
}

// ---- bulk-async staged inference sweep -----------------------------------------------------
// The direct kernel above walks a 4-deep dependent chain per slice (offset -> header -> records
// -> neighbour chain bytes) and spends ~540 warp instructions on a slice, a quarter of them on
// picking signs and on the hot/cold weight branch.  This kernel removes both costs:
//  * every warp owns a ring of `stages` slice buffers in shared memory and one mbarrier per
//    stage; lane 0 posts ONE 1-D bulk-async copy (cp.async.bulk global -> shared, the TMA unit's
//    non-tensor form, evict-first L2 policy) per slice -- the blob layout makes a slice one
//    contiguous 32 B-aligned run -- `stages` slices ahead of the one being sampled.  Slice
//    offsets come 32 at a time (lane = iteration) and are handed round by shuffle.  Headers and
//    records are then read with conflict-free LDS; what is left on the critical path of a slice
//    is shared-memory latency, the L2 gather of neighbour chain bytes (issued before the unary
//    records are summed) and arithmetic.  A slice longer than a stage (rare: the stage size is
//    picked to hold >= 99.5 % of the slices) is staged up to the stage's end and read from HBM.
//  * weights with id < n_tab are staged as a SIGNED table of three planes (-w | 0 | +w, n_tab
//    entries each): a record's sign code picks the plane, so a term costs two masks, one IMAD,
//    one LDS.64 and one DADD instead of compares and selects (bank conflicts as for a plain
//    table: the planes are a multiple of 128 B apart), and (ALLHOT) no hot/cold branch exists
//    at all when every weight fits.  (+-w) * fv == +-(w * fv) and x + (+-0) == x: the sums are
//    the direct kernel's bit for bit.
//  * the uniform of the draw (Philox) is computed and the neighbour chain bytes of the NEXT
//    slice (other colour classes: constant during the launch) are requested before the current
//    slice's records are summed, so the L2 gather is off the critical path too.
constexpr int SELL_MAX_STAGES = 4;
constexpr uint32_t SELL_MAX_TAB = 2048;   // weights in the signed table (3 planes x 8 B each)

__device__ __forceinline__ uint32_t sell_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sell_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void sell_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sell_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SELL_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SELL_DONE;\n"
      "bra SELL_WAIT;\n"
      "SELL_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t sell_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void sell_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}

// shared-memory plan of the staged kernel (host and device agree through these):
// signed weight table | one mbarrier per (warp, stage) | the rings
__host__ __device__ inline uint32_t sell_tab_plane(uint32_t n_tab) { return (n_tab * 8u + 127u) & ~127u; }  // bytes
__host__ __device__ inline uint32_t sell_staged_bar_off(uint32_t n_tab) { return 3u * sell_tab_plane(n_tab); }
__host__ __device__ inline uint32_t sell_staged_ring_off(uint32_t n_tab, uint32_t stages, uint32_t nwarp) {
  return (sell_staged_bar_off(n_tab) + nwarp * stages * 8u + 127u) & ~127u;
}
__host__ __device__ inline uint32_t sell_staged_smem(uint32_t n_tab, uint32_t stages, uint32_t stage_words,
                                                     uint32_t nwarp) {
  return sell_staged_ring_off(n_tab, stages, nwarp) + nwarp * stages * stage_words * 4u;
}

// the rare slice that is longer than a stage: potentials straight from HBM, as the direct kernel
// does (out of line, so that it does not cost the staged loop registers).  w_plain: no plain
// weight table is staged here, every weight comes through L1/L2 (n_hot = 0).
// (returns (pos, neg) by value: reference parameters would pin the caller's accumulators to local
// memory for the whole loop -- six STL.64 and two LDL.64 per slice in the round-1 kernel)
template <bool MULTI, int FV>
__device__ __noinline__ double2 sell_unstaged_potentials(const SellGraph &g, const SellState &st, uint32_t off,
                                                         uint32_t n_un, uint32_t n_mu, uint32_t flags, uint32_t lane,
                                                         uint32_t d) {
  double pos, neg;
  constexpr uint32_t FVW = FV == FV_F32 ? 1 : FV == FV_F64 ? 2 : 0;
  SellGraph gc = g;
  gc.n_hot = 0;
  SliceHeader h;
  h.n_un = n_un; h.n_mu = n_mu; h.n_oc = 0; h.evid = 0; h.pad = 0; h.evval = 0;
  h.flags = flags;
  h.un = g.blob + (size_t)off * 8 + SH_WORDS;
  h.mu = h.un + (size_t)32 * n_un * (1 + FVW);
  h.oth = h.mu + (size_t)32 * n_mu * (1 + FVW);
  sell_potentials<MULTI, FV, LD_STREAM>(gc, h, nullptr, st.weights, lane, d, st.assign_evid, pos, neg);
  return make_double2(pos, neg);
}

template <int FV>
__device__ __forceinline__ double sell_fv_s(const uint32_t *sm, uint32_t fvb, uint32_t j, uint32_t lane) {
  if (FV == FV_F32) return (double)__uint_as_float(sm[fvb + 32u * j + lane]);
  if (FV == FV_F64) return reinterpret_cast<const double *>(sm + fvb)[32u * j + lane];
  return 1.0;
}

// sign * weight: plane `code` (0 -> -w, 1 -> 0, 2 -> +w) of the signed table, wid8 = 8 * weight id
template <bool ALLHOT>
__device__ __forceinline__ double sell_signed_weight(const SellGraph &g, const unsigned char *tab_s, uint32_t plane,
                                                     const double *w, uint32_t code, uint32_t wid8) {
  if (ALLHOT) return *reinterpret_cast<const double *>(tab_s + (code * plane + wid8));
  // Branch-free: ONE generic load from either the signed table or the weight vector (L2), so that
  // the loads of a batch of records issue back to back; a branch per record would serialise a
  // warp's L2 round trips (nearly every record has some lane beyond the table).
  const bool hot = wid8 < 8u * g.n_tab;
  const double *ptr = hot ? reinterpret_cast<const double *>(tab_s + (code * plane + wid8)) : w + (wid8 >> 3);
  const double v = *ptr;
  return hot ? v : sell_term(code, v);
}

// one unary record of a staged slice into pos / neg (same arithmetic and order as sell_unary)
template <int FV, bool ANTISYM, bool ALLHOT>
__device__ __forceinline__ void sell_unary_rec(const SellGraph &g, const unsigned char *tab_s, uint32_t plane,
                                               const double *w, uint32_t rec, double fv, double &pos, double &neg) {
  const uint32_t wid8 = rec & SELL_UN_WID8_MASK;
  const double t1 = sell_signed_weight<ALLHOT>(g, tab_s, plane, w, rec & 3u, wid8);
  pos = __dadd_rn(pos, FV == FV_ONE ? t1 : __dmul_rn(t1, fv));
  if (!ANTISYM) {
    const double t0 = sell_signed_weight<ALLHOT>(g, tab_s, plane, w, rec >> 26, wid8);
    neg = __dadd_rn(neg, FV == FV_ONE ? t0 : __dmul_rn(t0, fv));
  }
}

template <int FV, bool ANTISYM, bool ALLHOT>
__device__ __forceinline__ void sell_unary_s(const SellGraph &g, const uint32_t *sm, uint32_t b, uint32_t n_un,
                                             const unsigned char *tab_s, uint32_t plane, const double *w,
                                             uint32_t lane, double &pos, double &neg) {
  // batches of 8 (table) / 10 (beyond the table: more weight loads in flight); the batch loop itself
  // is not unrolled -- the nest of remainder loops ptxas builds otherwise costs more than it saves
  // (C2: 0.243 -> 0.234 ms per sweep, profiles/r2_c2_ab.txt)
  constexpr uint32_t U = ALLHOT ? 8 : 10;
  const uint32_t un = b + SH_WORDS + lane, fvb = b + SH_WORDS + 32u * n_un;
  uint32_t j0 = 0;
#pragma unroll 1
  for (; j0 + U <= n_un; j0 += U) {
    uint32_t rec[U];
    double fv[U];
#pragma unroll
    for (uint32_t u = 0; u < U; ++u) {
      rec[u] = sm[un + 32u * (j0 + u)];
      fv[u] = sell_fv_s<FV>(sm, fvb, j0 + u, lane);
    }
#pragma unroll
    for (uint32_t u = 0; u < U; ++u) sell_unary_rec<FV, ANTISYM, ALLHOT>(g, tab_s, plane, w, rec[u], fv[u], pos, neg);
  }
  for (; j0 < n_un; ++j0)
    sell_unary_rec<FV, ANTISYM, ALLHOT>(g, tab_s, plane, w, sm[un + 32u * j0], sell_fv_s<FV>(sm, fvb, j0, lane), pos,
                                        neg);
  if (ANTISYM) neg = -pos;
}

constexpr uint32_t SELL_MC = 2;  // binary records handled together (4: no faster, profiles/r2_c2_ab.txt)

// wait for the slice in ring stage `b` (word index) and request the chain bytes of the other
// variables of its first SELL_MC binary records (raw bytes and equal_to bits: compared when the
// slice is sampled)
template <bool MULTI, int FV>
__device__ __forceinline__ void sell_staged_front(const SellState &st, const uint32_t *sm, uint32_t b, uint32_t bar,
                                                  uint32_t parity, uint32_t stage_words, uint32_t lane,
                                                  uint32_t (&av)[SELL_MC], uint32_t &eq) {
  constexpr uint32_t FVW = FV == FV_F32 ? 1 : FV == FV_F64 ? 2 : 0;
  sell_mbar_wait(bar, parity);
  if (!MULTI) return;
  const uint4 h0 = *reinterpret_cast<const uint4 *>(sm + b);
  const uint32_t flags = sm[b + SH_FLAGS];
  eq = 0;
  if ((flags & SF_BINARY) && SH_WORDS + 32u * ((h0.x + h0.y) * (1 + FVW) + h0.z) <= stage_words) {
    const uint32_t oth = b + SH_WORDS + 32u * (h0.x + h0.y) * (1 + FVW) + lane;
#pragma unroll
    for (uint32_t u = 0; u < SELL_MC; ++u) {
      const uint32_t ow = u < h0.y ? sm[oth + 32u * u] : 0u;
      av[u] = st.assign_evid[ow >> 1];  // (other colour classes: constant during the launch)
      eq |= (ow & 1u) << u;
    }
  }
}

// sample_single_variable (src/gibbs_sampler.h:151-169) for the 32 variables of the slice staged at
// word `b` of shared memory: potentials of both proposals, the Boolean draw with uniform r0, chain
// and tally.  av / eq: what sell_staged_front requested for this slice.
template <bool MULTI, int FV, bool ALLHOT, bool WIDE>
__device__ __forceinline__ void sell_staged_sample(const SellGraph &g, const SellState &st, const SweepParams &p,
                                                   const uint32_t *sm, const unsigned char *tab_s, uint32_t plane,
                                                   uint32_t b, uint32_t stage_words, uint32_t lane, uint32_t s,
                                                   double r0, uint32_t (&av)[SELL_MC], uint32_t eq) {
  constexpr uint32_t FVW = FV == FV_F32 ? 1 : FV == FV_F64 ? 2 : 0;
  constexpr uint32_t MC = SELL_MC;
  const uint32_t d = s * 32u + lane;
  const uint4 h0 = *reinterpret_cast<const uint4 *>(sm + b);      // n_un, n_mu, n_oc, evid
  const uint4 h1 = *reinterpret_cast<const uint4 *>(sm + b + 4);  // pad, evval, flags, own offset
  const uint32_t n_un = h0.x, n_mu = h0.y, flags = h1.z;
  // gibbs_sampler.h:157 -- a slice with nothing to sample is skipped
  const uint32_t active = ~h1.x & (p.sample_evidence ? 0xffffffffu : ~h0.w);
  if (!active) return;
  // FactorGraph::potential for proposals 1 and 0 (src/factor_graph.h:127-145)
  double pos = 0.0, neg = 0.0;
  if (SH_WORDS + 32u * ((n_un + n_mu) * (1 + FVW) + h0.z) > stage_words) {
    const double2 pn = sell_unstaged_potentials<MULTI, FV>(g, st, h1.w, n_un, n_mu, flags, lane, sell_self_slot<WIDE>(s, flags, lane));
    pos = pn.x; neg = pn.y;
  } else {
    const uint32_t mu = b + SH_WORDS + 32u * n_un * (1 + FVW) + lane;
    const uint32_t mufv = b + SH_WORDS + 32u * n_un * (1 + FVW) + 32u * n_mu;
    const uint32_t oth = mu + 32u * n_mu * (1 + FVW);
    if (flags & SF_ANTISYM) sell_unary_s<FV, true, ALLHOT>(g, sm, b, n_un, tab_s, plane, st.weights, lane, pos, neg);
    else sell_unary_s<FV, false, ALLHOT>(g, sm, b, n_un, tab_s, plane, st.weights, lane, pos, neg);
    if (MULTI && (flags & SF_BINARY)) {
      // record j owns others word j; its tag is the factor's sign table code(x, sat) at bits
      // 2 (2 x + sat): x = 1 feeds pos, x = 0 feeds neg
      for (uint32_t j0 = 0; j0 < n_mu; j0 += MC) {
        if (j0) {
          uint32_t ow[MC];
          eq = 0;
#pragma unroll
          for (uint32_t u = 0; u < MC; ++u) ow[u] = j0 + u < n_mu ? sm[oth + 32u * (j0 + u)] : 0u;
#pragma unroll
          for (uint32_t u = 0; u < MC; ++u) av[u] = st.assign_evid[ow[u] >> 1];
#pragma unroll
          for (uint32_t u = 0; u < MC; ++u) eq |= (ow[u] & 1u) << u;
        }
#pragma unroll
        for (uint32_t u = 0; u < MC; ++u) {
          const bool ok = j0 + u < n_mu;
          const uint32_t rec = ok ? sm[mu + 32u * (j0 + u)] : (SELL_BINARY_NOP << 24);
          const uint32_t sat = av[u] == ((eq >> u) & 1u) ? 1u : 0u;
          const uint32_t codes = rec >> (24u + 2u * sat), wid8 = (rec & SELL_WID_MASK) << 3;
          double t1 = sell_signed_weight<ALLHOT>(g, tab_s, plane, st.weights, (codes >> 4) & 3u, wid8);
          double t0 = sell_signed_weight<ALLHOT>(g, tab_s, plane, st.weights, codes & 3u, wid8);
          if (FV != FV_ONE) {
            const double fv = ok ? sell_fv_s<FV>(sm, mufv, j0 + u, lane) : 0.0;
            t1 = __dmul_rn(t1, fv);
            t0 = __dmul_rn(t0, fv);
          }
          pos = __dadd_rn(pos, t1);
          neg = __dadd_rn(neg, t0);
        }
      }
    } else if (MULTI && n_mu) {
      SellGraph gc = g;
      gc.n_hot = 0;  // no plain weight table here: the general path reads weights through L1/L2
      SliceHeader h;
      h.n_un = n_un; h.n_mu = n_mu; h.n_oc = h0.z; h.evid = h0.w; h.pad = h1.x; h.evval = 0;
      h.flags = flags;
      h.un = sm + b + SH_WORDS; h.mu = sm + mu - lane; h.oth = sm + oth - lane;
      sell_general_multi<FV, LD_SMEM>(gc, h, nullptr, st.weights, lane, sell_self_slot<WIDE>(s, flags, lane), st.assign_evid, pos, neg);
    }
  }
  if (WIDE && (flags & SF_WIDE)) {  // warp per variable (hubs)
    sell_wide_draw(st, p, s, flags, lane, active, pos, neg);
  } else if ((active >> lane) & 1u) {
    const uint32_t prop = sell_draw_boolean(pos, neg, r0);
    st.assign_evid[d] = (uint8_t)prop;
    if (prop) atomicAdd(st.tallies + d, 1u);  // gibbs_sampler.h:163-167
  }
}

// signed weight table, one mbarrier per (warp, stage)
__device__ __forceinline__ void sell_staged_prologue(const SellGraph &g, const SellState &st, unsigned char *smem,
                                                     uint32_t plane, uint32_t bar0, uint32_t stages, uint32_t lane) {
  if (lane == 0)
    for (uint32_t i = 0; i < stages; ++i) sell_mbar_init(bar0 + 8u * i, 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (uint32_t i = threadIdx.x; i < g.n_tab; i += blockDim.x) {
    const double w = __ldg(st.weights + i);
    *reinterpret_cast<double *>(smem + 8u * i) = -w;
    *reinterpret_cast<double *>(smem + plane + 8u * i) = 0.0;
    *reinterpret_cast<double *>(smem + 2u * plane + 8u * i) = w;
  }
  __syncthreads();
}

template <bool MULTI, int FV, bool ALLHOT, int NT, bool WIDE>
__global__ void __launch_bounds__(NT, 2)
k_sell_infer_staged(const SellGraph g, const SellState st, const SweepParams p, const uint32_t stages,
                    const uint32_t stage_words) {
  extern __shared__ __align__(128) unsigned char sell_smem[];
  constexpr uint32_t MC = SELL_MC;
  // Programmatic dependent launch: the launch of the next colour class may start as soon as SMs
  // free up, run its prologue (weight table, mbarriers, slice offsets, the first bulk copies --
  // all of it data no sweep writes) and park at griddepcontrol.wait below until this class is
  // complete and visible.  Fills the tail of one class with the ramp of the next.
  asm volatile("griddepcontrol.launch_dependents;");
  const unsigned char *tab_s = sell_smem;
  const uint32_t plane = sell_tab_plane(g.n_tab);
  const uint32_t *sm = reinterpret_cast<const uint32_t *>(sell_smem);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t bar0 = sell_smem_addr(sell_smem) + sell_staged_bar_off(g.n_tab) + warp * stages * 8u;
  const uint32_t ringw = sell_staged_ring_off(g.n_tab, stages, NT / 32) / 4u + warp * stages * stage_words;  // word index
  const uint32_t ring0 = sell_smem_addr(sell_smem) + ringw * 4u;
  sell_staged_prologue(g, st, sell_smem, plane, bar0, stages, lane);

  // slices of this warp: s(k) = s0 + k * step, k < n_my
  const uint64_t step = (uint64_t)((gridDim.x * blockDim.x) >> 5) * g.slice_stride;
  const uint64_t s0 = g.slice_begin + (uint64_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * g.slice_stride;
  if (s0 >= g.slice_end) return;
  const uint32_t n_my = (uint32_t)((g.slice_end - 1 - s0) / step) + 1u;
  const uint64_t pol = sell_policy_evict_first();

  // slice offsets, 32 iterations per fetch (lane = iteration), one block ahead
  uint32_t cur_lo = 0, cur_hi = 0, nxt_lo = 0, nxt_hi = 0;
  auto fetch_offsets = [&](uint32_t kb, uint32_t &lo, uint32_t &hi) {
    const uint32_t k = kb + lane;
    if (k < n_my) {
      const uint32_t *o = g.sl_off + (s0 + (uint64_t)k * step);
      lo = __ldg(o);
      hi = __ldg(o + 1);
    }
  };
  fetch_offsets(0, cur_lo, cur_hi);
  fetch_offsets(32, nxt_lo, nxt_hi);
  // post the copy of iteration k into stage `stage` (called for k = 0, 1, 2, ... in order)
  auto issue = [&](uint32_t k, uint32_t stage) {
    const uint32_t j = k & 31u;
    if (j == 0 && k) {
      cur_lo = nxt_lo; cur_hi = nxt_hi;
      fetch_offsets(k + 32, nxt_lo, nxt_hi);
    }
    const uint32_t off = __shfl_sync(0xffffffffu, cur_lo, j);
    const uint32_t bytes = min((__shfl_sync(0xffffffffu, cur_hi, j) - off) * 32u, stage_words * 4u);
    if (lane == 0) {
      sell_mbar_expect_tx(bar0 + 8u * stage, bytes);
      sell_bulk_g2s(ring0 + stage * stage_words * 4u, g.blob + (size_t)off * 8, bytes, bar0 + 8u * stage, pol);
    }
  };

  const uint32_t pre = n_my < stages ? n_my : stages;
  for (uint32_t k = 0; k < pre; ++k) issue(k, k);
  uint32_t stage = 0, parity = 0;  // of iteration k
  uint32_t av[MC] = {}, eq = 0, nav[MC] = {}, neq = 0;
  // from here on the kernel reads and writes the chains and tallies: the previous class must be done
  asm volatile("griddepcontrol.wait;" ::: "memory");
  sell_staged_front<MULTI, FV>(st, sm, ringw, bar0, 0, stage_words, lane, av, eq);
  uint64_t s64 = s0;
  for (uint32_t k = 0; k < n_my; ++k, s64 += step) {
    uint32_t nstage = stage + 1, nparity = parity;
    if (nstage == stages) { nstage = 0; nparity ^= 1u; }
    double r0, r1;
    dev_uniforms(p, (uint32_t)s64 * 32u + lane, r0, r1);
    if (k + 1 < n_my) sell_staged_front<MULTI, FV>(st, sm, ringw + nstage * stage_words, bar0 + 8u * nstage, nparity, stage_words, lane, nav, neq);
    sell_staged_sample<MULTI, FV, ALLHOT, WIDE>(g, st, p, sm, tab_s, plane, ringw + stage * stage_words, stage_words, lane, (uint32_t)s64, r0, av, eq);
    __syncwarp();  // every lane is done with this stage before the next copy lands in it
    if (k + stages < n_my) issue(k + stages, stage);
    stage = nstage; parity = nparity;
#pragma unroll
    for (uint32_t u = 0; u < MC; ++u) av[u] = nav[u];
    eq = neq;
  }
}

// ---- learning sweep (sample_sgd, src/gibbs_sampler.cc:71-75) --------------------------------
// The reference's SGD is per sample: every variable moves its weights at once
// (src/inference_result.h:66-85) and the next variable of the id-order scan draws with the moved
// weights.  Two things follow (tests/sim/): (1) the learned weights weight the
// variables visited LAST most (relaxation time 1/(stepsize x curvature) updates), so a sweep must
// visit the variables in ascending id like the reference -- SellLearn::visit, slices in ascending
// id block; (2) only about cap/stepsize updates of a weight may be in flight before its step has
// to shrink to stay stable.  One kernel, launched twice per epoch by the host:
//   wide  every SM, the head of the visit order: fast, steps of heavily shared weights scaled down
//         (s_k = min(1, cap / (stepsize x in-flight curvature of k))); brings every weight near its
//         fixed point and keeps both chains moving;
//   tail  the last few relaxation times of the visit order on as few warps as the cap allows with
//         s_k = 1: the reference's own recursion up to a Hogwild staleness of a few dozen updates
//         (its -t N threads have N).
// A CTA keeps a copy of the hot weights in shared memory and refreshes it from the live global
// vector every `period` slices per warp: draws of a period see the weights as of its start,
// gradients of the period are pre-summed in shared memory and pushed with one atomic per weight.
struct SellLearn {
  const uint32_t *visit;      // slice indices in learning order
  uint32_t vbegin, vend;      // positions of this launch
  uint32_t period;            // slices per warp between two refreshes of the CTA's weight copy
  double step;                // stepsize of this epoch
  double xcap_inv;            // (stepsize x in-flight fraction x coupling) / cap: s_k = min(1, 1 / (xcap_inv * H_k))
  double decay;               // reg * step (l2) or reg (l1, sic: no stepsize, src/inference_result.h:77)
  double inv_nslice;          // 1 / slices of the whole visit order
  int reg_l1;
  const float *w_N, *w_H;     // per weight: SGD updates / curvature bound of one whole sweep
  double *weights;            // live
};

// step scale of weight k (1 = the reference's step)
__device__ __forceinline__ double sell_step_scale(const SellLearn &L, uint32_t wid) {
  const double x = L.xcap_inv * (double)__ldg(L.w_H + wid);
  return x > 1.0 ? 1.0 / x : 1.0;
}

// gradient of one SGD step: hot weights into the CTA's accumulator of the period, others straight
// into the live vector (update_weight, src/inference_result.h:66-85).  The DECAY of the weights
// outside the shared copy is applied per launch by k_cold_decay from their static update counts --
// like the hot weights' decay (lumped per period) it does not ride on the records: a record whose
// gradient is zero (the free draw agrees with the evidence) costs no atomic.  C4 (1 M weights, 23 %
// of the record visits outside the copy) went from 24.3 to the figure in DESIGN.md 4 with it.
// Graphs whose feature values are all 1 have integer gradients (differences of sign codes): their
// accumulators are int32 (a native shared-memory atomic instead of a compare-and-swap loop).
template <int FV>
__device__ __forceinline__ void sell_grad(const SellGraph &g, const SellLearn &L, double *acc_s, uint32_t wid,
                                          bool fixed, double grad) {
  if (fixed) return;  // factor_graph.cc:247
  if (wid < g.n_hot) {
    if (grad == 0.0) return;
    if (FV == FV_ONE) atomicAdd(reinterpret_cast<int *>(acc_s + wid), (int)grad);
    else atomicAdd(acc_s + wid, grad);
    return;
  }
  if (grad == 0.0) return;
  atomicAdd(L.weights + wid, -L.step * sell_step_scale(L, wid) * grad);
}

// decay of the weights outside the shared-memory copy for one learning launch that visited `share`
// of the sweep: weight k took N_k * share SGD steps, each  w <- w / (1 + reg step)  resp.
// w += reg (w < 0)  (src/inference_result.h:72-80)
__global__ void k_cold_decay(const SellLearn L, const uint8_t *fixed, uint32_t n_hot, uint32_t nweight, double share) {
  const uint32_t i = n_hot + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nweight || fixed[i]) return;
  const double n = (double)L.w_N[i] * share;
  if (n == 0.0) return;
  const double x = L.weights[i];
  L.weights[i] = L.reg_l1 ? x + n * L.decay * (x < 0 ? 1.0 : 0.0) : x * exp(-n * log1p(L.decay * sell_step_scale(L, i)));
}
template <int FV>
__device__ __forceinline__ double sell_acc_take(double *acc_s, uint32_t i) {  // read and clear
  double a;
  if (FV == FV_ONE) { a = (double)*reinterpret_cast<int *>(acc_s + i); }
  else a = acc_s[i];
  acc_s[i] = 0.0;  // all-zero bits: 0 in both views
  return a;
}

// The update of weight i after a learning launch (see k_apply_weights below for the scheme).
__device__ __forceinline__ void sell_apply_weight(uint32_t i, double *w, double *grad_acc, const float *N,
                                                  const float *Q, const float *H, const uint8_t *fixed, double step,
                                                  double reg, int reg_l1, int decay_in_acc, double share) {
  if (fixed[i]) return;
  const double nn = (double)N[i] * share;  // this launch swept `share` of the class
  const double acc = __ldcg(grad_acc + i);  // summed by atomics in L2
  if (nn == 0.0 && acc == 0.0) return;
  grad_acc[i] = 0.0;
  const double x = w[i];
  const double r = reg_l1 ? 0.0 : reg * step / (1.0 + reg * step);
  double S = acc;
  if (!decay_in_acc) S += reg_l1 ? -nn * reg * (x < 0 ? 1.0 : 0.0) / step : nn * (r / step) * x;
  // curvature D: the weight's own (diagonal) bound H where the joint step is provably
  // non-expansive with it, i.e. step * phi(step * D) * R <= 1 with R the row-sum majorizer;
  // otherwise the smallest D that is (fixed point of D = R (1 - exp(-step D)), approached from above)
  const double Rk = (double)Q[i] * share, Hk = (double)H[i] * share;
  double D = Hk;
  if (Rk * -expm1(-step * D) > D) {
    D = Rk;
    for (int it = 0; it < 12; ++it) D = Rk * -expm1(-step * D);
  }
  const double z = step * D + nn * r;
  const double phi = z > 1e-8 ? -expm1(-z) / z : 1.0 - 0.5 * z;
  w[i] = x - step * S * phi;
}

// sgd_on_variable (factor_graph.cc:262-277) over the records of slice h held by this lane:
// evidence_value = assignment_dense (ev), free-chain value pf.
// grad = sign_free * fv - sign_evid * fv == (code_free - code_evid) * fv exactly.
template <bool MULTI, int FV, int LD>
__device__ __forceinline__ void sell_sgd_records(const SellGraph &g, const SellState &st, const SellLearn &L,
                                                 const SliceHeader &h, double *acc_s, uint32_t lane, uint32_t d,
                                                 uint32_t pf, uint32_t ev, bool learn, bool masks, uint32_t sf_mask,
                                                 uint32_t se_mask) {
  if (__any_sync(0xffffffffu, learn && pf != ev)) {  // unary: same value on both chains = zero gradient
    // records (and feature values) in batches: a slice longer than a ring stage is read from L2 / HBM
    // here, and one record per trip exposed a full round trip per record (C4: a quarter of the stalls)
    constexpr uint32_t U = FV == FV_F64 ? 4 : 8;
    for (uint32_t j0 = 0; j0 < h.n_un; j0 += U) {
      uint32_t rec[U];
      double fv[U];
#pragma unroll
      for (uint32_t u = 0; u < U; ++u) {
        const bool ok = j0 + u < h.n_un;
        rec[u] = ok ? sell_ld<LD>(h.un + (size_t)32 * (j0 + u) + lane) : SELL_UNARY_NOP;
        fv[u] = (FV != FV_ONE && ok) ? sell_fv<FV, LD>(h.un, h.n_un, j0 + u, lane) : 1.0;
      }
#pragma unroll
      for (uint32_t u = 0; u < U; ++u) {
        if (!learn || rec[u] == SELL_UNARY_NOP) continue;
        const int cf = (int)(pf ? rec[u] & 3u : rec[u] >> 26), ce = (int)(ev ? rec[u] & 3u : rec[u] >> 26);
        if (cf == ce) continue;
        sell_grad<FV>(g, L, acc_s, sell_unary_wid(rec[u]), rec[u] & SELL_UN_FIXED_BIT, __dmul_rn((double)(cf - ce), fv[u]));
      }
    }
  }
  if (!MULTI || !h.n_mu) return;
  if (h.flags & SF_BINARY) {
    for (uint32_t j = 0; j < h.n_mu; ++j) {
      const uint32_t rec = sell_ld<LD>(h.mu + (size_t)32 * j + lane);
      const uint32_t ow = sell_ld<LD>(h.oth + (size_t)32 * j + lane);
      if (!learn) continue;
      const uint32_t tag = rec >> 24, od = ow >> 1, eq = ow & 1u;
      uint32_t sf, se;
      if (masks) { sf = (sf_mask >> j) & 1u; se = (se_mask >> j) & 1u; }
      else { sf = sell_chain<LD>(st.assign_free, od) == eq; se = sell_chain<LD>(st.assign_evid, od) == eq; }
      const int cf = (int)((tag >> (4u * pf + 2u * sf)) & 3u), ce = (int)((tag >> (4u * ev + 2u * se)) & 3u);
      if (cf == ce || tag == SELL_BINARY_NOP) continue;
      sell_grad<FV>(g, L, acc_s, rec & SELL_WID_MASK, rec & SELL_FIXED_BIT,
                __dmul_rn((double)(cf - ce), sell_fv<FV, LD>(h.mu, h.n_mu, j, lane)));
    }
  } else {
    uint32_t oc = 0;
    for (uint32_t j = 0; j < h.n_mu; ++j) {
      const uint32_t rec = sell_ld<LD>(h.mu + (size_t)32 * j + lane);
      const uint32_t tag = rec >> 24;
      if ((tag >> TAG_FORM_SHIFT) == 0) continue;
      if (!learn) { oc += sell_general_span<LD>(h, tag, lane, oc); continue; }
      uint32_t oc2 = oc;
      double f0, f1, e0, e1;
      sell_general_signs<LD>(h, tag, d, lane, oc, st.assign_free, f0, f1);
      sell_general_signs<LD>(h, tag, d, lane, oc2, st.assign_evid, e0, e1);
      const double fv = sell_fv<FV, LD>(h.mu, h.n_mu, j, lane);
      sell_grad<FV>(g, L, acc_s, rec & SELL_WID_MASK, rec & SELL_FIXED_BIT,
                __dsub_rn(__dmul_rn(pf ? f1 : f0, fv), __dmul_rn(ev ? e1 : e0, fv)));
    }
  }
}

// One variable per lane of one slice of a learning sweep (sample_sgd_single_variable,
// gibbs_sampler.h:127-149): both chains, then the SGD gradients of the lane's records into the
// launch's accumulators.  LD: where the slice is read from (HBM through L1, or a staged copy).
template <bool MULTI, int FV, int LD, bool WIDE>
__device__ __forceinline__ void sell_learn_slice(const SellGraph &g, const SellState &st, const SweepParams &p,
                                                 const SellLearn &L, const SliceHeader &h, const double *w_s,
                                                 double *acc_s, uint32_t lane, uint32_t d) {
  // SF_WIDE: the slice is ONE variable (slot 0) whose records are dealt to the 32 lanes -- every lane
  // takes the variable's flags, uniforms and draws, sums its share of the records and pushes the SGD
  // steps of its share; lane 0 writes the chains
  const bool wide = WIDE && (h.flags & SF_WIDE);
  const uint32_t hl = wide ? 0u : lane;
  if (wide) d -= lane;
  const bool writer = !wide || lane == 0;
  const bool pad = (h.pad >> hl) & 1u, evid = (h.evid >> hl) & 1u;
  const uint32_t ev = (h.evval >> hl) & 1u;  // Variable::assignment_dense
  double r0, r1;
  dev_uniforms(p, d, r0, r1);
  // Slices of two-variable factors (the common coupled case): the chain values of every other
  // variable, both chains, are gathered in ONE batch of independent L2 loads and kept as two bit
  // masks (bit j: the other variable of record j satisfies its equal_to) -- the free-chain
  // potentials, the evid-chain potentials and the SGD pass below all read the masks.
  const bool masks = MULTI && (h.flags & SF_BINARY) && h.n_mu <= 32u;
  uint32_t sf_mask = 0, se_mask = 0;
  if (masks) {
    for (uint32_t j0 = 0; j0 < h.n_mu; j0 += 4) {
      uint32_t ow[4], af[4], ae[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) ow[u] = j0 + u < h.n_mu ? sell_ld<LD>(h.oth + (size_t)32 * (j0 + u) + lane) : 0u;
#pragma unroll
      for (int u = 0; u < 4; ++u) { af[u] = sell_chain<LD>(st.assign_free, ow[u] >> 1); ae[u] = sell_chain<LD>(st.assign_evid, ow[u] >> 1); }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + u < h.n_mu) {
          sf_mask |= (af[u] == (ow[u] & 1u) ? 1u : 0u) << (j0 + u);
          se_mask |= (ae[u] == (ow[u] & 1u) ? 1u : 0u) << (j0 + u);
        }
    }
  }
  // the multi-variable records of a mask slice into (pos, neg) for chain mask m
  auto masked_multi = [&](uint32_t m, double &ps, double &ng) {
    for (uint32_t j = 0; j < h.n_mu; ++j) {
      const uint32_t rec = sell_ld<LD>(h.mu + (size_t)32 * j + lane);
      const double wt = sell_weight<LD>(g, w_s, st.weights, rec & SELL_WID_MASK);
      const double t = FV == FV_ONE ? wt : __dmul_rn(wt, sell_fv<FV, LD>(h.mu, h.n_mu, j, lane));
      const uint32_t tab = (rec >> 24) >> (2u * ((m >> j) & 1u));
      ps = __dadd_rn(ps, sell_term((tab >> 4) & 3u, t));
      ng = __dadd_rn(ng, sell_term(tab & 3u, t));
    }
  };
  // evid chain (sample_evid, gibbs_sampler.h:171-190): clamp hard evidence, else a Gibbs draw
  // (Boolean variables carry no truthiness, so --noise_aware makes evidence variables draw too)
  const uint32_t draw_evid = wide ? ((!pad && (p.noise_aware || !evid)) ? 0xffffffffu : 0u)
                                  : ~h.pad & (p.noise_aware ? 0xffffffffu : ~h.evid);
  const bool part = WIDE && (h.flags & SF_PART);  // a part of a hub (graph_image.h): see sell_part_finish
  const uint32_t sfirst = d / 32u - (part ? (h.flags >> 4) & 0x3fffu : 0u);
  uint32_t pf, pe = ev;
  {
    // free chain: every variable, evidence included (gibbs_sampler.h:137-139)
    double pos, neg, upos, uneg;
    if (masks) {
      pos = 0.0; neg = 0.0;
      if (h.flags & SF_ANTISYM) sell_unary<FV, LD, true>(g, h, w_s, st.weights, lane, pos, neg);
      else sell_unary<FV, LD, false>(g, h, w_s, st.weights, lane, pos, neg);
      upos = pos; uneg = neg;
      masked_multi(sf_mask, pos, neg);
    } else {
      sell_potentials<MULTI, FV, LD>(g, h, w_s, st.weights, lane, part ? sfirst * 32u : d, st.assign_free, pos, neg, &upos, &uneg);
    }
    double epos = pos, eneg = neg;  // unary-only slices: both chains see identical potentials
    if (draw_evid && MULTI && h.n_mu) {
      epos = upos; eneg = uneg;  // the unary sums are the same on both chains: continue from them
      if (masks) masked_multi(se_mask, epos, eneg);
      else sell_multi<FV, LD>(g, h, w_s, st.weights, lane, part ? sfirst * 32u : d, st.assign_evid, nullptr, 0u, epos, eneg);
    }
    if (wide) { pos = sell_wide_sum(pos); neg = sell_wide_sum(neg); epos = sell_wide_sum(epos); eneg = sell_wide_sum(eneg); }
    if (part) {
      double v[4] = {pos, neg, epos, eneg};
      if (!sell_part_finish<4>(st, d / 32u, h.flags, lane, v)) return;  // not the hub's last part
      pos = v[0]; neg = v[1]; epos = v[2]; eneg = v[3];
      d = sfirst * 32u;
      dev_uniforms(p, d, r0, r1);
    }
    pf = sell_draw_boolean(pos, neg, r0);
    if (draw_evid && ((draw_evid >> lane) & 1u)) pe = sell_draw_boolean(epos, eneg, r1);
    if (!pad && writer) {
      st.assign_free[d] = (uint8_t)pf;
      st.assign_evid[d] = (uint8_t)pe;
    }
  }
  // which variables learn (gibbs_sampler.h:144-146); Boolean has_truthiness() is false
  const bool learn = !pad && (p.learn_non_evidence || (!p.noise_aware && evid));
  if (!__any_sync(0xffffffffu, learn)) return;
  if (!part) {
    sell_sgd_records<MULTI, FV, LD>(g, st, L, h, acc_s, lane, d, pf, ev, learn, masks, sf_mask, se_mask);
    return;
  }
  // a hub: this warp finished it and drew for it; it now pushes the SGD steps of ALL its parts, read
  // from the blob (the other parts were summed by other warps), at the hub's place in the visit order
  const uint32_t nparts = h.flags >> 18;
  for (uint32_t q = 0; q < nparts; ++q) {
    const SliceHeader hq = sell_header<FV>(g, sfirst + q);
    sell_sgd_records<MULTI, FV, LD_CACHED>(g, st, L, hq, acc_s, lane, d, pf, ev, learn, false, 0u, 0u);
  }
}

// the CTA's copy of the hot weights from the live vector (L2), four loads in flight per thread
__device__ __forceinline__ void sell_refresh_weights(double *w_s, const double *w, uint32_t n_hot) {
  for (uint32_t i0 = threadIdx.x; i0 < n_hot; i0 += 4u * blockDim.x) {
    double v[4];
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u) v[u] = i0 + u * blockDim.x < n_hot ? __ldcg(w + i0 + u * blockDim.x) : 0.0;
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u)
      if (i0 + u * blockDim.x < n_hot) w_s[i0 + u * blockDim.x] = v[u];
  }
}

// One launch of a learning sweep: positions [vbegin, vend) of the visit order, dealt to the warps
// of the grid round-robin (position vbegin + round * warps + warp: the whole grid moves through the
// order together, so "later in the order" is "later in time").  Slices come through a per-warp
// two-stage ring of bulk-async copies like the staged inference kernel (a slice longer than a
// stage is read from HBM instead): the blob is static, so the ring runs ahead across the periods.
constexpr int SELL_LEARN_THREADS = 256;
constexpr uint32_t SELL_LEARN_STAGES = 2;
__host__ __device__ inline uint32_t sell_learn_bar_off(uint32_t n_hot) { return (28u * n_hot + 15u) & ~15u; }
__host__ __device__ inline uint32_t sell_learn_ring_off(uint32_t n_hot) {
  return (sell_learn_bar_off(n_hot) + (SELL_LEARN_THREADS / 32) * SELL_LEARN_STAGES * 8u + 127u) & ~127u;
}
__host__ __device__ inline size_t sell_learn_smem(uint32_t n_hot, uint32_t stage_words) {
  return sell_learn_ring_off(n_hot) + (size_t)(SELL_LEARN_THREADS / 32) * SELL_LEARN_STAGES * stage_words * 4u;
}

template <bool MULTI, int FV, bool WIDE>
__global__ void __launch_bounds__(SELL_LEARN_THREADS, 2)
k_sell_learn_live(const SellGraph g, const SellState st, const SweepParams p, const SellLearn L,
                  const uint32_t stage_words) {
  extern __shared__ __align__(128) unsigned char sell_smem[];
  constexpr uint32_t FVW = FV == FV_F32 ? 1 : FV == FV_F64 ? 2 : 0;
  constexpr uint32_t ST = SELL_LEARN_STAGES;
  double *w_s = reinterpret_cast<double *>(sell_smem), *acc_s = w_s + g.n_hot;
  float *as_s = reinterpret_cast<float *>(acc_s + g.n_hot), *lam_s = as_s + g.n_hot, *debt_s = lam_s + g.n_hot;
  const uint32_t *sm = reinterpret_cast<const uint32_t *>(sell_smem);
  __shared__ uint32_t cnt_s;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t npos = L.vend - L.vbegin, rounds = (npos + warps - 1) / warps;
  const uint32_t n_my = gw < npos ? (npos - 1 - gw) / warps + 1u : 0u;  // positions of this warp
  const uint32_t bar0 = sell_smem_addr(sell_smem) + sell_learn_bar_off(g.n_hot) + warp * ST * 8u;
  const uint32_t ringw = sell_learn_ring_off(g.n_hot) / 4u + warp * ST * stage_words;  // word index
  const uint32_t ring0 = sell_smem_addr(sell_smem) + ringw * 4u;
  if (lane == 0)
    for (uint32_t i = 0; i < ST; ++i) sell_mbar_init(bar0 + 8u * i, 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // per hot weight: step of one update (stepsize x scale) and its decay per slice of the sweep
  // (N_k / slices updates per slice, each  w <- w / (1 + reg step)  resp.  w += reg (w < 0))
  for (uint32_t i = threadIdx.x; i < g.n_hot; i += blockDim.x) {
    const double sc = sell_step_scale(L, i);
    as_s[i] = (float)(L.step * sc);
    const double per_update = L.reg_l1 ? L.decay : log1p(L.decay * sc);  // (1 + reg step)^-n = exp(-n log1p(reg step))
    lam_s[i] = (float)((double)__ldg(L.w_N + i) * L.inv_nslice * per_update);
    acc_s[i] = 0.0;
    debt_s[i] = 0.f;
  }
  // slice index and offsets of this warp's positions, 32 at a time (lane = iteration), one block ahead
  const uint64_t pol = sell_policy_evict_first();
  uint32_t cur_s = 0, cur_lo = 0, cur_hi = 0, nxt_s = 0, nxt_lo = 0, nxt_hi = 0;
  auto fetch = [&](uint32_t kb, uint32_t &sl, uint32_t &lo, uint32_t &hi) {
    const uint32_t k = kb + lane;
    if (k < n_my) {
      sl = __ldg(L.visit + L.vbegin + (size_t)k * warps + gw);
      lo = __ldg(g.sl_off + sl);
      hi = __ldg(g.sl_off + sl + 1);
    }
  };
  auto issue = [&](uint32_t k, uint32_t stage) {  // called for k = 0, 1, 2, ... in order
    const uint32_t j = k & 31u;
    if (j == 0 && k) {
      cur_s = nxt_s; cur_lo = nxt_lo; cur_hi = nxt_hi;
      fetch(k + 32, nxt_s, nxt_lo, nxt_hi);
    }
    const uint32_t off = __shfl_sync(0xffffffffu, cur_lo, j);
    const uint32_t bytes = min((__shfl_sync(0xffffffffu, cur_hi, j) - off) * 32u, stage_words * 4u);
    if (lane == 0) {
      sell_mbar_expect_tx(bar0 + 8u * stage, bytes);
      sell_bulk_g2s(ring0 + stage * stage_words * 4u, g.blob + (size_t)off * 8, bytes, bar0 + 8u * stage, pol);
    }
  };
  __syncthreads();  // mbarriers initialised before any lane waits on one
  if (n_my) {
    fetch(0, cur_s, cur_lo, cur_hi);
    fetch(32, nxt_s, nxt_lo, nxt_hi);
    const uint32_t pre = n_my < ST ? n_my : ST;
    for (uint32_t k = 0; k < pre; ++k) issue(k, k);
  }
  // the slice index of iteration k was fetched with block k / 32: keep the block of the iteration
  // being PROCESSED beside the one being issued (which runs ST ahead)
  uint32_t stage = 0, parity = 0, k = 0;
  uint32_t proc_s = cur_s;  // slice indices of the block being processed
  for (uint32_t r0 = 0; r0 < rounds; r0 = r0 + L.period < r0 ? rounds : r0 + L.period) {
    __syncthreads();  // the previous flush is done with w_s / acc_s
    sell_refresh_weights(w_s, L.weights, g.n_hot);
    if (threadIdx.x == 0) cnt_s = 0;
    __syncthreads();
    uint32_t mine = 0;
    const uint32_t r1 = r0 + L.period < r0 ? rounds : min(rounds, r0 + L.period);
    for (uint32_t r = r0; r < r1 && k < n_my; ++r, ++k) {
      // slice index of iteration k: issue(k) ran two iterations ago and moved cur_s on to k's block
      // when k opens one; the two last iterations of a block still read the block's own copy
      if ((k & 31u) == 0 && k) proc_s = cur_s;
      const uint32_t s = __shfl_sync(0xffffffffu, proc_s, k & 31u);
      sell_mbar_wait(bar0 + 8u * stage, parity);
      const uint32_t *b = sm + ringw + stage * stage_words;
      const uint4 h0 = *reinterpret_cast<const uint4 *>(b);
      const uint4 h1 = *reinterpret_cast<const uint4 *>(b + 4);
      SliceHeader h;
      h.n_un = h0.x; h.n_mu = h0.y; h.n_oc = h0.z; h.evid = h0.w;
      h.pad = h1.x; h.evval = h1.y; h.flags = h1.z;
      if (SH_WORDS + 32u * ((h.n_un + h.n_mu) * (1 + FVW) + h.n_oc) <= stage_words) {
        h.un = b + SH_WORDS;
        h.mu = h.un + 32u * h.n_un * (1 + FVW);
        h.oth = h.mu + 32u * h.n_mu * (1 + FVW);
        sell_learn_slice<MULTI, FV, LD_SMEM_LIVE, WIDE>(g, st, p, L, h, w_s, acc_s, lane, s * 32u + lane);
      } else {  // longer than a stage: the records come from HBM (the header carries the slice's offset)
        h.un = g.blob + (size_t)h1.w * 8 + SH_WORDS;
        h.mu = h.un + (size_t)32 * h.n_un * (1 + FVW);
        h.oth = h.mu + (size_t)32 * h.n_mu * (1 + FVW);
        sell_learn_slice<MULTI, FV, LD_CACHED, WIDE>(g, st, p, L, h, w_s, acc_s, lane, s * 32u + lane);
      }
      ++mine;
      __syncwarp();  // every lane is done with this stage before the next copy lands in it
      if (k + ST < n_my) issue(k + ST, stage);
      if (++stage == ST) { stage = 0; parity ^= 1u; }
    }
    // slices this CTA swept in the period (for the decay of the updates it made)
    if (lane == 0 && mine) atomicAdd(&cnt_s, mine);
    __syncthreads();
    const uint32_t cnt = cnt_s;
    // push the period's steps.  The decay of a weight nobody touched in the period is carried as a
    // debt and paid with the next step of the weight (or once it reaches 1e-3, or at the end of the
    // launch): a CTA holding a thousand idle weights does not send a thousand atomics per period.
    const bool last = r1 >= rounds;
    for (uint32_t i = threadIdx.x; i < g.n_hot; i += blockDim.x) {
      const double a = sell_acc_take<FV>(acc_s, i), x = w_s[i];
      const float debt = debt_s[i] + lam_s[i] * (float)cnt;
      if (a == 0.0 && debt < 1e-3f && !last) { debt_s[i] = debt; continue; }
      debt_s[i] = 0.f;
      const double dec = L.reg_l1 ? (double)debt * (x < 0 ? 1.0 : 0.0) : x * expm1(-(double)debt);
      const double d = dec - (double)as_s[i] * a;
      if (d != 0.0) atomicAdd(L.weights + i, d);
    }
  }
}

// The weight change of one learning launch.  The reference applies, per SGD step i on weight
// w (src/inference_result.h:66-85):   w <- w / (1 + reg*a) - a * g_i      (l2, a = stepsize)
// A launch makes N such steps with the weights frozen, so it integrates them at once:
//     S  = sum g_i  +  N * reg/(1+reg*a) * w          (gradient + decay of the N steps)
//     z  = a * D  +  N * reg*a/(1+reg*a)               (D >= curvature of the N steps, see below)
//     w <- w - a * S * (1 - exp(-z)) / z
// For z -> 0 this is exactly the sum of the reference's N steps; for large z (a weight shared
// by thousands of concurrent updates, where N plain steps from a stale weight would overshoot)
// it tends to the damped Newton step -S/(Q + N*reg), i.e. the same fixed point
// E[g] + reg*w = 0 the sequential recursion relaxes to, reached without oscillation.
// l1 (sic, `w += reg * (w < 0)` per step): the decay term is  -N * reg * (w < 0) / a.
__global__ void k_apply_weights(double *w, double *grad_acc, const float *N, const float *Q, const float *H,
                                const uint8_t *fixed, uint32_t n, double step, double reg, int reg_l1,
                                int decay_in_acc, double share) {
  // chained by programmatic dependent launch between two learning launches: lets the next one
  // start its prologue, then waits for the gradient sums of the previous one
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sell_apply_weight(i, w, grad_acc, N, Q, H, fixed, step, reg, reg_l1, decay_in_acc, share);
}

}  // namespace dwg
#endif
