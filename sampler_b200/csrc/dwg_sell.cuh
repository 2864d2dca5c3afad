// dwg_sell.cuh -- the production sweep kernels for Boolean graphs (sm_100a): one warp per
// 32-variable slice of the sliced-ELL stream (graph_image.h), lane = variable.
//
//   k_sell_infer   GibbsSamplerThread::sample        (src/gibbs_sampler.cc:65-69,
//                                                      src/gibbs_sampler.h:151-169,192-215)
//   k_sell_learn   GibbsSamplerThread::sample_sgd    (src/gibbs_sampler.cc:71-75,
//                  src/gibbs_sampler.h:127-149,171-190; sgd_on_variable/-factor,
//                  src/factor_graph.cc:243-314)
//   k_apply_weights  the weight update of one learning launch (update_weight,
//                  src/inference_result.h:66-85, as a semi-implicit mini-batch step)
//
// Memory behaviour (HBM-bound by design): every edge record is read exactly once per sweep
// with warp-wide contiguous requests (32 B of tags, 128 B of weight ids, 128 B of feature
// values per column); per variable 1 B flags, 1 B chain write, 8 B tally read-modify-write.
// Gathers that stay on chip: weights (ids < n_hot from shared memory, the rest through L1/L2)
// and the u8 chain values of other variables (L2).  Persistent CTAs (a few per SM) walk the
// slices of a colour class with a grid stride so the shared-memory weight table is staged once
// per CTA.  No tensor cores: O(1) flop per 5..13 B of edge data.
#ifndef DWG_SELL_CUH_
#define DWG_SELL_CUH_
#include "dwg_device.cuh"

namespace dwg {

constexpr int SELL_THREADS = 512;  // 16 warps = 16 slices in flight per CTA

struct SellGraph {
  uint32_t slice_begin, slice_end;  // slices of this launch (one colour class)
  const uint32_t *sl_col;           // [nslice+1]
  const uint32_t *sl_ocol;          // [nslice+1] or null
  const uint8_t *tag;
  const uint32_t *wid;
  const float *fv32;
  const double *fv64;
  const uint32_t *oth;
  const uint8_t *var_flags;
  const uint8_t *evid_val;  // u8 chains only (Boolean)
  const uint8_t *w_fixed;   // internal weight order
  uint32_t nweight, n_hot;
};

struct SellState {
  uint8_t *assign_free, *assign_evid;
  const double *weights;  // frozen for the duration of a launch
  double *grad_acc;       // [nweight] sum of gradients of this launch (learning)
  uint32_t *tallies;
};

template <int FV>
__device__ __forceinline__ double sell_fv(const SellGraph &g, size_t idx) {
  if (FV == FV_F32) return (double)__ldg(g.fv32 + idx);
  if (FV == FV_F64) return __ldg(g.fv64 + idx);
  return 1.0;
}

// sign code (0,1,2) -> the term  sign * t  with t = weight * feature value
__device__ __forceinline__ double sell_term(uint32_t code, double t) {
  return code == 2u ? t : (code == 0u ? -t : 0.0);
}
__device__ __forceinline__ double sell_code_sign(uint32_t code) { return (double)((int)code - 1); }

// One multi-variable edge record of lane's variable: signs when the variable is 0 / 1, other
// variables read from chain `a`.  Consumes this lane's words of the others stream
// (word c of the lane at (ocol0 + c) * 32 + lane).
__device__ __forceinline__ void sell_multi_signs(const SellGraph &g, uint32_t tag, uint32_t self,
                                                 size_t obase, uint32_t &oc, const uint8_t *a,
                                                 double &s0, double &s1) {
  const uint32_t fc = tag & TAG_FUNC_MASK, form = tag >> TAG_FORM_SHIFT;
  if (form != FORM_GENERAL) {
    const uint32_t self_eq = (tag & TAG_SELF_EQ1) ? 1u : 0u;
    const bool self_head = tag & TAG_SELF_HEAD;
    uint32_t nsb = 0;
    bool head_oth = false;
    for (uint32_t j = 0; j < form; ++j) {
      const uint32_t w = __ldg(g.oth + obase + (size_t)(oc + j) * 32);
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
  const uint32_t n = __ldg(g.oth + obase + (size_t)oc * 32);
  uint32_t nsb0 = 0, nsb1 = 0;
  bool h0 = false, h1 = false;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t w = __ldg(g.oth + obase + (size_t)(oc + 1 + i) * 32);
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

// number of others words behind a multi record (to skip it)
__device__ __forceinline__ uint32_t sell_multi_span(const SellGraph &g, uint32_t tag, size_t obase, uint32_t oc) {
  const uint32_t form = tag >> TAG_FORM_SHIFT;
  return form != FORM_GENERAL ? form : 1u + __ldg(g.oth + obase + (size_t)oc * 32);
}

__device__ __forceinline__ double sell_weight(const SellGraph &g, const double *w_s, const double *w, uint32_t wid) {
  return wid < g.n_hot ? w_s[wid] : __ldg(w + wid);
}

__device__ __forceinline__ void sell_stage_weights(const SellGraph &g, const double *w, double *w_s) {
  for (uint32_t i = threadIdx.x; i < g.n_hot; i += blockDim.x) w_s[i] = __ldg(w + i);
}

// FactorGraph::potential for proposals 1 and 0 of the lane's variable, in index order
// (src/factor_graph.h:127-145): pos/neg += weight * (sign * feature_value)
template <bool MULTI, int FV>
__device__ __forceinline__ void sell_potentials(const SellGraph &g, const double *w_s, const double *w,
                                                uint32_t col0, uint32_t ncol, size_t obase, uint32_t lane,
                                                uint32_t self, const uint8_t *a, double &pos, double &neg) {
  pos = 0.0; neg = 0.0;
  uint32_t oc = 0;
#pragma unroll 2
  for (uint32_t j = 0; j < ncol; ++j) {
    const size_t idx = (size_t)(col0 + j) * 32 + lane;
    const uint32_t tag = __ldg(g.tag + idx);
    if (tag == SELL_NOP) continue;
    const uint32_t wid = __ldg(g.wid + idx);
    const double wt = sell_weight(g, w_s, w, wid);
    const double fv = sell_fv<FV>(g, idx);
    if (!MULTI || (tag >> TAG_FORM_SHIFT) == 0) {
      const double t = FV == FV_ONE ? wt : __dmul_rn(wt, fv);  // w * (+-1 * fv) == +-(w * fv) exactly
      pos = __dadd_rn(pos, sell_term(tag & 3u, t));
      neg = __dadd_rn(neg, sell_term((tag >> 2) & 3u, t));
    } else {
      double s0, s1;
      sell_multi_signs(g, tag, self, obase, oc, a, s0, s1);
      pos = dev_acc(pos, wt, s1, fv);
      neg = dev_acc(neg, wt, s0, fv);
    }
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
  for (uint32_t s = g.slice_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); s < g.slice_end; s += warps) {
    const uint32_t d = s * 32u + lane;
    const uint32_t fl = __ldg(g.var_flags + d);
    const bool active = !(fl & VF_PAD) && (!(fl & VF_EVID) || p.sample_evidence);  // gibbs_sampler.h:157
    // slices are sorted by evidence: a slice with nothing to sample is skipped as a whole
    if (!__any_sync(0xffffffffu, active)) continue;
    const uint32_t col0 = __ldg(g.sl_col + s), ncol = __ldg(g.sl_col + s + 1) - col0;
    const size_t obase = MULTI ? (size_t)__ldg(g.sl_ocol + s) * 32 + lane : 0;
    double pos, neg;
    sell_potentials<MULTI, FV>(g, w_s, st.weights, col0, ncol, obase, lane, d, st.assign_evid, pos, neg);
    if (!active) continue;
    double r0, r1;
    dev_uniforms(p, d, r0, r1);
    const uint32_t prop = dev_draw_boolean(pos, neg, r0);
    st.assign_evid[d] = (uint8_t)prop;
    if (prop) st.tallies[d] += 1u;  // gibbs_sampler.h:163-167 (agg_nsamples is implicit)
  }
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
  for (uint32_t s = g.slice_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); s < g.slice_end; s += warps) {
    const uint32_t d = s * 32u + lane;
    const uint32_t fl = __ldg(g.var_flags + d);
    const bool pad = fl & VF_PAD, evid = fl & VF_EVID;
    const uint32_t col0 = __ldg(g.sl_col + s), ncol = __ldg(g.sl_col + s + 1) - col0;
    const size_t obase = MULTI ? (size_t)__ldg(g.sl_ocol + s) * 32 + lane : 0;
    double r0, r1;
    dev_uniforms(p, d, r0, r1);
    // free chain: every variable, evidence included (gibbs_sampler.h:137-139)
    double pos, neg;
    sell_potentials<MULTI, FV>(g, w_s, st.weights, col0, ncol, obase, lane, d, st.assign_free, pos, neg);
    const uint32_t pf = dev_draw_boolean(pos, neg, r0);
    if (!pad) st.assign_free[d] = (uint8_t)pf;
    // evid chain (sample_evid, gibbs_sampler.h:171-190): clamp hard evidence, else a Gibbs draw
    // (Boolean variables carry no truthiness, so --noise_aware makes evidence variables draw too)
    const uint32_t ev = __ldg(g.evid_val + d);
    uint32_t pe = ev;
    const bool draw_evid = !pad && (p.noise_aware || !evid);
    if (__any_sync(0xffffffffu, draw_evid)) {
      double pe_pos = pos, pe_neg = neg;
      if (MULTI)  // unary-only graphs: both chains see identical potentials
        sell_potentials<MULTI, FV>(g, w_s, st.weights, col0, ncol, obase, lane, d, st.assign_evid, pe_pos, pe_neg);
      if (draw_evid) pe = dev_draw_boolean(pe_pos, pe_neg, r1);
    }
    if (!pad) st.assign_evid[d] = (uint8_t)pe;
    // which variables learn (gibbs_sampler.h:144-146); Boolean has_truthiness() is false
    const bool learn = !pad && (p.learn_non_evidence || (!p.noise_aware && evid));
    if (!__any_sync(0xffffffffu, learn)) continue;
    // sgd_on_variable (factor_graph.cc:262-277): evidence_value = assignment_dense, free = pf
    uint32_t oc = 0;
    for (uint32_t j = 0; j < ncol; ++j) {
      const size_t idx = (size_t)(col0 + j) * 32 + lane;
      const uint32_t tag = __ldg(g.tag + idx);
      if (tag == SELL_NOP) continue;
      const uint32_t wid = __ldg(g.wid + idx);
      double grad;
      if (!MULTI || (tag >> TAG_FORM_SHIFT) == 0) {
        if (!learn || pf == ev) continue;  // same sign on both chains: zero gradient
        const double sf = sell_code_sign(pf ? (tag & 3u) : ((tag >> 2) & 3u));
        const double se = sell_code_sign(ev ? (tag & 3u) : ((tag >> 2) & 3u));
        grad = __dmul_rn(__dsub_rn(sf, se), sell_fv<FV>(g, idx));
      } else {
        if (!learn) { oc += sell_multi_span(g, tag, obase, oc); continue; }
        uint32_t oc2 = oc;
        double f0, f1, e0, e1;
        sell_multi_signs(g, tag, d, obase, oc, st.assign_free, f0, f1);
        sell_multi_signs(g, tag, d, obase, oc2, st.assign_evid, e0, e1);
        const double fv = sell_fv<FV>(g, idx);
        grad = __dsub_rn(__dmul_rn(pf ? f1 : f0, fv), __dmul_rn(ev ? e1 : e0, fv));
      }
      if (grad == 0.0 || __ldg(g.w_fixed + wid)) continue;  // factor_graph.cc:247
      if (wid < g.n_hot) atomicAdd(acc_s + wid, grad);
      else atomicAdd(st.grad_acc + wid, grad);
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
//     z  = a * Q  +  N * reg*a/(1+reg*a)               (Q >= sum of the steps' curvature)
//     w <- w - a * S * (1 - exp(-z)) / z
// For z -> 0 this is exactly the sum of the reference's N steps; for large z (a weight shared
// by thousands of concurrent updates, where N plain steps from a stale weight would overshoot)
// it tends to the damped Newton step -S/(Q + N*reg), i.e. the same fixed point
// E[g] + reg*w = 0 the sequential recursion relaxes to, reached without oscillation.
// l1 (sic, `w += reg * (w < 0)` per step): the decay term is  -N * reg * (w < 0) / a.
__global__ void k_apply_weights(double *w, double *grad_acc, const float *N, const float *Q,
                                const uint8_t *fixed, uint32_t n, double step, double reg, int reg_l1,
                                int decay_in_acc) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || fixed[i]) return;
  const double nn = (double)N[i];
  const double acc = grad_acc[i];
  if (nn == 0.0 && acc == 0.0) return;
  grad_acc[i] = 0.0;
  const double x = w[i];
  const double r = reg_l1 ? 0.0 : reg * step / (1.0 + reg * step);
  double S = acc;
  if (!decay_in_acc) S += reg_l1 ? -nn * reg * (x < 0 ? 1.0 : 0.0) / step : nn * (r / step) * x;
  const double z = step * (double)Q[i] + nn * r;
  const double phi = z > 1e-8 ? -expm1(-z) / z : 1.0 - 0.5 * z;
  w[i] = x - step * S * phi;
}

}  // namespace dwg
#endif
