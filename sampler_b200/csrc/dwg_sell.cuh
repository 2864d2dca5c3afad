// dwg_sell.cuh -- the production sweep kernels for Boolean graphs (sm_100a): one warp per
// 32-variable slice of the sliced-ELL blob (graph_image.h), lane = variable.
//
//   k_sell_infer   GibbsSamplerThread::sample        (src/gibbs_sampler.cc:65-69,
//                                                      src/gibbs_sampler.h:151-169,192-215)
//   k_sell_learn   GibbsSamplerThread::sample_sgd    (src/gibbs_sampler.cc:71-75,
//                  src/gibbs_sampler.h:127-149,171-190; sgd_on_variable/-factor,
//                  src/factor_graph.cc:243-314)
//   k_apply_weights  the weight update of one learning launch (update_weight,
//                  src/inference_result.h:66-85, as a semi-implicit mini-batch step)
//
// Memory behaviour (HBM-bound by design): a slice is one contiguous run of words, read exactly
// once per sweep with warp-wide 128 B requests; all record loads of a slice (unary chunk,
// multi chunk, others words) are issued before the first dependent use, so every warp keeps
// ~1.5 KB in flight and an SM (32 warps) ~50 KB -- what 6.5 TB/s x ~1 us of latency needs.
// Per variable besides the records: 1 B chain write, 8 B tally read-modify-write; evidence /
// padding / evidence-value bits ride in the slice header.  Gathers that stay on chip: weights
// (ids < n_hot from shared memory, the rest through L1/L2) and the u8 chain values of other
// variables (L2).  The fast paths are branch-free: a unary record carries its two sign codes,
// a record of a two-variable factor carries the factor's whole 2x2 sign table, so no
// factor-function switch runs per edge.  Persistent CTAs (2 per SM) walk the slices of a colour
// class with a grid stride so the shared-memory weight table is staged once per CTA.
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
};

struct SellState {
  uint8_t *assign_free, *assign_evid;
  const double *weights;  // frozen for the duration of a launch
  double *grad_acc;       // [nweight] sum of gradients of this launch (learning)
  uint32_t *tallies;
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

// streaming (evict-first) loads for the inference sweep, cached ones for learning (its SGD
// pass walks the slice a second time and finds it in L1)
template <bool STREAM>
__device__ __forceinline__ uint32_t sell_ld(const uint32_t *p) { return STREAM ? __ldcs(p) : __ldg(p); }

// feature value of record (column j, lane) of a section whose records start at `sec` and
// hold `ncol` columns
template <int FV, bool STREAM>
__device__ __forceinline__ double sell_fv(const uint32_t *sec, uint32_t ncol, uint32_t j, uint32_t lane) {
  if (FV == FV_F32) {
    const float *p = reinterpret_cast<const float *>(sec + (size_t)32 * ncol) + (size_t)32 * j + lane;
    return (double)(STREAM ? __ldcs(p) : __ldg(p));
  }
  if (FV == FV_F64) {
    const double *p = reinterpret_cast<const double *>(sec + (size_t)32 * ncol) + (size_t)32 * j + lane;
    return STREAM ? __ldcs(p) : __ldg(p);
  }
  return 1.0;
}

// sign code (0,1,2) -> the term  sign * t  with t = weight * feature value
__device__ __forceinline__ double sell_term(uint32_t code, double t) {
  return code == 2u ? t : (code == 0u ? -t : 0.0);
}

__device__ __forceinline__ double sell_weight(const SellGraph &g, const double *w_s, const double *w, uint32_t wid) {
  return wid < g.n_hot ? w_s[wid] : __ldg(w + wid);
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
template <int FV, bool STREAM, bool ANTISYM>
__device__ __forceinline__ void sell_unary(const SellGraph &g, const SliceHeader &h, const double *w_s,
                                           const double *w, uint32_t lane, double &pos, double &neg) {
  constexpr int U = FV == FV_F64 ? 4 : 8;
  for (uint32_t j0 = 0; j0 < h.n_un; j0 += U) {
    uint32_t rec[U];
    double fv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = j0 + u < h.n_un;
      rec[u] = ok ? sell_ld<STREAM>(h.un + (size_t)32 * (j0 + u) + lane) : (SELL_UNARY_NOP << 24);
      fv[u] = (FV != FV_ONE && ok) ? sell_fv<FV, STREAM>(h.un, h.n_un, j0 + u, lane) : 1.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double wt = sell_weight(g, w_s, w, rec[u] & SELL_WID_MASK);
      const double t = FV == FV_ONE ? wt : __dmul_rn(wt, fv[u]);  // w * (+-1 * fv) == +-(w * fv) exactly
      pos = __dadd_rn(pos, sell_term((rec[u] >> 24) & 3u, t));
      if (!ANTISYM) neg = __dadd_rn(neg, sell_term((rec[u] >> 26) & 3u, t));
    }
  }
  if (ANTISYM) neg = -pos;  // round-to-nearest is sign-symmetric: the mirrored sum, bit for bit
}

// ---- multi records of a general slice: CSR tag semantics, per-lane cursor into the others ----
// Signs of one record when the variable is 0 / 1, the other variables read from chain `a`.
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
      const uint32_t w = __ldg(o + (size_t)(oc + j) * 32);
      const bool s = (uint32_t)a[w >> 1] == (w & 1u);
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
  const uint32_t n = __ldg(o + (size_t)oc * 32);
  uint32_t nsb0 = 0, nsb1 = 0;
  bool h0 = false, h1 = false;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t w = __ldg(o + (size_t)(oc + 1 + i) * 32);
    const uint32_t d = w >> 1, eq = w & 1u;
    bool t0, t1;
    if (d == self) { t0 = eq == 0u; t1 = eq == 1u; }
    else { t0 = t1 = (uint32_t)a[d] == eq; }
    if (i == n - 1) { h0 = t0; h1 = t1; }
    else { nsb0 += t0; nsb1 += t1; }
  }
  oc += 1 + n;
  s0 = dev_sign(fc, n, nsb0, h0);
  s1 = dev_sign(fc, n, nsb1, h1);
}

// number of others words behind a general-slice record (to skip it)
__device__ __forceinline__ uint32_t sell_general_span(const SliceHeader &h, uint32_t tag, uint32_t lane, uint32_t oc) {
  const uint32_t form = tag >> TAG_FORM_SHIFT;
  return form != FORM_GENERAL ? form : 1u + __ldg(h.oth + (size_t)oc * 32 + lane);
}

template <int FV, bool STREAM>
__device__ __forceinline__ void sell_general_multi(const SellGraph &g, const SliceHeader &h, const double *w_s,
                                                   const double *w, uint32_t lane, uint32_t self,
                                                   const uint8_t *a, double &pos, double &neg) {
  uint32_t oc = 0;
  for (uint32_t j = 0; j < h.n_mu; ++j) {
    const uint32_t rec = sell_ld<STREAM>(h.mu + (size_t)32 * j + lane);
    const uint32_t tag = rec >> 24;
    if ((tag >> TAG_FORM_SHIFT) == 0) continue;  // padding
    const double wt = sell_weight(g, w_s, w, rec & SELL_WID_MASK);
    const double fv = sell_fv<FV, STREAM>(h.mu, h.n_mu, j, lane);
    double s0, s1;
    sell_general_signs(h, tag, self, lane, oc, a, s0, s1);
    pos = dev_acc(pos, wt, s1, fv);
    neg = dev_acc(neg, wt, s0, fv);
  }
}

// ---- multi records of a binary slice: record j owns others word j; the tag is the factor's
// sign table over (value of this variable, other variable satisfied) ----
struct BinaryChunk {
  uint32_t rec[SELL_MCHUNK], ow[SELL_MCHUNK];
};

template <bool STREAM>
__device__ __forceinline__ void sell_binary_load(const SliceHeader &h, uint32_t j0, uint32_t lane, BinaryChunk &c) {
#pragma unroll
  for (int u = 0; u < SELL_MCHUNK; ++u) {
    const bool ok = j0 + u < h.n_mu;
    c.rec[u] = ok ? sell_ld<STREAM>(h.mu + (size_t)32 * (j0 + u) + lane) : (SELL_BINARY_NOP << 24);
    c.ow[u] = ok ? sell_ld<STREAM>(h.oth + (size_t)32 * (j0 + u) + lane) : 0u;
  }
}

// bit u: the other variable of record u satisfies its equal_to on chain `a`
__device__ __forceinline__ uint32_t sell_binary_sat(const BinaryChunk &c, const uint8_t *a) {
  uint32_t av[SELL_MCHUNK], sat = 0;
#pragma unroll
  for (int u = 0; u < SELL_MCHUNK; ++u) av[u] = a[c.ow[u] >> 1];
#pragma unroll
  for (int u = 0; u < SELL_MCHUNK; ++u) sat |= (av[u] == (c.ow[u] & 1u) ? 1u : 0u) << u;
  return sat;
}

template <int FV, bool STREAM>
__device__ __forceinline__ void sell_binary_acc(const SellGraph &g, const SliceHeader &h, const double *w_s,
                                                const double *w, uint32_t j0, uint32_t lane,
                                                const BinaryChunk &c, uint32_t sat, double &pos, double &neg) {
#pragma unroll
  for (int u = 0; u < SELL_MCHUNK; ++u) {
    const double wt = sell_weight(g, w_s, w, c.rec[u] & SELL_WID_MASK);
    double t = wt;
    if (FV != FV_ONE) t = __dmul_rn(wt, j0 + u < h.n_mu ? sell_fv<FV, STREAM>(h.mu, h.n_mu, j0 + u, lane) : 0.0);
    const uint32_t tab = (c.rec[u] >> 24) >> (2u * ((sat >> u) & 1u));
    pos = __dadd_rn(pos, sell_term((tab >> 4) & 3u, t));
    neg = __dadd_rn(neg, sell_term(tab & 3u, t));
  }
}

// FactorGraph::potential for proposals 1 and 0 of the lane's variable on chain `a`
// (src/factor_graph.h:127-145); `first` = the slice's first binary chunk, loaded by the caller
// ahead of the unary records so that its gathers overlap them
template <bool MULTI, int FV, bool STREAM>
__device__ __forceinline__ void sell_potentials(const SellGraph &g, const SliceHeader &h, const double *w_s,
                                                const double *w, uint32_t lane, uint32_t self,
                                                const uint8_t *a, double &pos, double &neg) {
  pos = 0.0; neg = 0.0;
  BinaryChunk c;
  uint32_t sat = 0;
  const bool binary = MULTI && (h.flags & SF_BINARY);
  if (binary) {
    sell_binary_load<STREAM>(h, 0, lane, c);
    sat = sell_binary_sat(c, a);
  }
  if (h.flags & SF_ANTISYM) sell_unary<FV, STREAM, true>(g, h, w_s, w, lane, pos, neg);
  else sell_unary<FV, STREAM, false>(g, h, w_s, w, lane, pos, neg);
  if (!MULTI) return;
  if (binary) {
    sell_binary_acc<FV, STREAM>(g, h, w_s, w, 0, lane, c, sat, pos, neg);
    for (uint32_t j0 = SELL_MCHUNK; j0 < h.n_mu; j0 += SELL_MCHUNK) {
      sell_binary_load<STREAM>(h, j0, lane, c);
      sat = sell_binary_sat(c, a);
      sell_binary_acc<FV, STREAM>(g, h, w_s, w, j0, lane, c, sat, pos, neg);
    }
  } else if (h.n_mu) {
    sell_general_multi<FV, STREAM>(g, h, w_s, w, lane, self, a, pos, neg);
  }
}

template <bool MULTI, int FV>
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
    sell_potentials<MULTI, FV, true>(g, h, w_s, st.weights, lane, d, st.assign_evid, pos, neg);
    if (!((active >> lane) & 1u)) continue;
    double r0, r1;
    dev_uniforms(p, d, r0, r1);
    const uint32_t prop = sell_draw_boolean(pos, neg, r0);
    st.assign_evid[d] = (uint8_t)prop;
    if (prop) st.tallies[d] += 1u;  // gibbs_sampler.h:163-167 (agg_nsamples is implicit)
  }
}

// gradient of one SGD step into the launch's accumulators (hot weights: shared memory)
__device__ __forceinline__ void sell_grad(const SellGraph &g, const SellState &st, double *acc_s, uint32_t rec,
                                          double grad) {
  if (grad == 0.0 || (rec & SELL_FIXED_BIT)) return;  // factor_graph.cc:247
  const uint32_t wid = rec & SELL_WID_MASK;
  if (wid < g.n_hot) atomicAdd(acc_s + wid, grad);
  else atomicAdd(st.grad_acc + wid, grad);
}

// One learning sweep over a colour class.  Weights are FROZEN for the launch; the gradient
// of every SGD step (pot_free - pot_evid, src/factor_graph.cc:255-258) is summed per weight
// into grad_acc (hot weights pre-aggregated in shared memory), and k_apply_weights turns the
// sums into the weight change afterwards.
template <bool MULTI, int FV>
__global__ void __launch_bounds__(SELL_THREADS, 2)
k_sell_learn(const SellGraph g, const SellState st, const SweepParams p) {
  extern __shared__ double smem[];
  double *w_s = smem, *acc_s = smem + g.n_hot;
  sell_stage_weights(g, st.weights, w_s);
  for (uint32_t i = threadIdx.x; i < g.n_hot; i += blockDim.x) acc_s[i] = 0.0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint64_t s64 = g.slice_begin + (uint64_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * g.slice_stride;
       s64 < g.slice_end; s64 += (uint64_t)warps * g.slice_stride) {
    const uint32_t s = (uint32_t)s64;
    const SliceHeader h = sell_header<FV>(g, s);
    const uint32_t d = s * 32u + lane;
    const bool pad = (h.pad >> lane) & 1u, evid = (h.evid >> lane) & 1u;
    const uint32_t ev = (h.evval >> lane) & 1u;  // Variable::assignment_dense
    double r0, r1;
    dev_uniforms(p, d, r0, r1);
    // free chain: every variable, evidence included (gibbs_sampler.h:137-139)
    double pos, neg;
    sell_potentials<MULTI, FV, false>(g, h, w_s, st.weights, lane, d, st.assign_free, pos, neg);
    const uint32_t pf = sell_draw_boolean(pos, neg, r0);
    if (!pad) st.assign_free[d] = (uint8_t)pf;
    // evid chain (sample_evid, gibbs_sampler.h:171-190): clamp hard evidence, else a Gibbs draw
    // (Boolean variables carry no truthiness, so --noise_aware makes evidence variables draw too)
    uint32_t pe = ev;
    const uint32_t draw_evid = ~h.pad & (p.noise_aware ? 0xffffffffu : ~h.evid);
    if (draw_evid) {
      if (MULTI && h.n_mu)  // unary-only slices: both chains see identical potentials
        sell_potentials<MULTI, FV, false>(g, h, w_s, st.weights, lane, d, st.assign_evid, pos, neg);
      if ((draw_evid >> lane) & 1u) pe = sell_draw_boolean(pos, neg, r1);
    }
    if (!pad) st.assign_evid[d] = (uint8_t)pe;
    // which variables learn (gibbs_sampler.h:144-146); Boolean has_truthiness() is false
    const bool learn = !pad && (p.learn_non_evidence || (!p.noise_aware && evid));
    if (!__any_sync(0xffffffffu, learn)) continue;
    // sgd_on_variable (factor_graph.cc:262-277): evidence_value = assignment_dense, free = pf.
    // grad = sign_free * fv - sign_evid * fv == (code_free - code_evid) * fv exactly.
    if (__any_sync(0xffffffffu, learn && pf != ev)) {  // unary: same value on both chains = zero gradient
      for (uint32_t j = 0; j < h.n_un; ++j) {
        const uint32_t rec = __ldg(h.un + (size_t)32 * j + lane);
        if (!learn || pf == ev) continue;
        const uint32_t tag = rec >> 24;
        const int cf = (int)((tag >> (pf ? 0 : 2)) & 3u), ce = (int)((tag >> (ev ? 0 : 2)) & 3u);
        if (cf == ce) continue;
        sell_grad(g, st, acc_s, rec, __dmul_rn((double)(cf - ce), sell_fv<FV, false>(h.un, h.n_un, j, lane)));
      }
    }
    if (!MULTI || !h.n_mu) continue;
    if (h.flags & SF_BINARY) {
      for (uint32_t j = 0; j < h.n_mu; ++j) {
        const uint32_t rec = __ldg(h.mu + (size_t)32 * j + lane);
        const uint32_t ow = __ldg(h.oth + (size_t)32 * j + lane);
        if (!learn) continue;
        const uint32_t tag = rec >> 24, od = ow >> 1, eq = ow & 1u;
        const uint32_t sf = st.assign_free[od] == eq, se = st.assign_evid[od] == eq;
        const int cf = (int)((tag >> (4u * pf + 2u * sf)) & 3u), ce = (int)((tag >> (4u * ev + 2u * se)) & 3u);
        if (cf == ce) continue;
        sell_grad(g, st, acc_s, rec, __dmul_rn((double)(cf - ce), sell_fv<FV, false>(h.mu, h.n_mu, j, lane)));
      }
    } else {
      uint32_t oc = 0;
      for (uint32_t j = 0; j < h.n_mu; ++j) {
        const uint32_t rec = __ldg(h.mu + (size_t)32 * j + lane);
        const uint32_t tag = rec >> 24;
        if ((tag >> TAG_FORM_SHIFT) == 0) continue;
        if (!learn) { oc += sell_general_span(h, tag, lane, oc); continue; }
        uint32_t oc2 = oc;
        double f0, f1, e0, e1;
        sell_general_signs(h, tag, d, lane, oc, st.assign_free, f0, f1);
        sell_general_signs(h, tag, d, lane, oc2, st.assign_evid, e0, e1);
        const double fv = sell_fv<FV, false>(h.mu, h.n_mu, j, lane);
        sell_grad(g, st, acc_s, rec, __dsub_rn(__dmul_rn(pf ? f1 : f0, fv), __dmul_rn(ev ? e1 : e0, fv)));
      }
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < g.n_hot; i += blockDim.x) {
    const double a = acc_s[i];
    if (a != 0.0) atomicAdd(st.grad_acc + i, a);
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
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || fixed[i]) return;
  const double nn = (double)N[i] * share;  // this launch swept `share` of the class
  const double acc = grad_acc[i];
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

}  // namespace dwg
#endif
