// dwg_kernels.cuh -- the sweep kernels (sm_100a): thread-per-variable Gibbs sweeps over
// the schedule-ordered CSR (graph_image.h).
//
//   k_sweep_infer  replaces GibbsSamplerThread::sample -> sample_single_variable
//                  (src/gibbs_sampler.cc:65-69, src/gibbs_sampler.h:151-169)
//   k_sweep_learn  replaces GibbsSamplerThread::sample_sgd -> sample_sgd_single_variable
//                  (src/gibbs_sampler.cc:71-75, src/gibbs_sampler.h:127-149) with
//                  sgd_on_variable / sgd_on_factor (src/factor_graph.cc:243-314) and
//                  update_weight (src/inference_result.h:66-85) as atomic updates
// One launch covers one colour class [d_begin, d_end).  HBM-bound: the edge arrays are
// streamed once per sweep; weights and the u8 chains are L2-resident gathers.
#ifndef DWG_KERNELS_CUH_
#define DWG_KERNELS_CUH_
#include "dwg_device.cuh"

namespace dwg {

constexpr int SWEEP_THREADS = 256;

template <typename AssignT, bool LIVE>
__device__ __forceinline__ uint32_t dev_draw_sample(const DevGraph &g, uint32_t d, bool cat,
                                                    uint32_t row0, uint32_t K, const AssignT *a,
                                                    const double *w, double r) {
  if (!cat) {
    double pos, neg;
    dev_bool_potentials<AssignT, LIVE>(g, d, row0, a, w, pos, neg);
    return dev_draw_boolean(pos, neg, r);
  }
  return dev_draw_categorical<AssignT, LIVE>(g, d, row0, K, a, w, r);
}

// sample_single_variable (src/gibbs_sampler.h:151-169) for the variable at schedule position d
template <typename AssignT>
__device__ __forceinline__ void sweep_infer_var(const DevGraph &g, const DevState &st, const SweepParams &p, uint32_t d) {
  AssignT *a = static_cast<AssignT *>(st.assign_evid);
  const uint8_t fl = __ldg(g.var_flags + d);
  if (fl & VF_PAD) return;
  if ((fl & VF_EVID) && !p.sample_evidence) return;  // gibbs_sampler.h:157
  const bool cat = fl & VF_CATEGORICAL;
  uint32_t row0 = d, K = 1;
  if (g.var_row) {
    row0 = __ldg(g.var_row + d);
    K = __ldg(g.var_row + d + 1) - row0;
  }
  double r0, r1;
  dev_uniforms(p, d, r0, r1);
  const uint32_t prop = dev_draw_sample<AssignT, false>(g, d, cat, row0, K, a, st.weights, r0);
  a[d] = (AssignT)prop;  // :160
  // agg_nsamples is implicit (every sampled variable gains one per sweep, :163);
  // Boolean: tally only proposal == 1 at the single row; categorical: row of the value (:164-167)
  // (a RED: nothing waits for the old count)
  if (cat) atomicAdd(st.tallies + row0 + prop, 1u);
  else if (prop == 1u) atomicAdd(st.tallies + row0, 1u);
}

template <typename AssignT>
__global__ void __launch_bounds__(SWEEP_THREADS)
k_sweep_infer(const DevGraph g, const DevState st, const SweepParams p) {
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;; tid += nthreads) {
    const uint64_t d64 = sweep_unit_var(p, tid);
    if (d64 >= p.d_end) {  // past the class: done once the whole 32-variable unit is
      if (d64 - (tid & 31u) >= p.d_end) break;
      continue;
    }
    sweep_infer_var<AssignT>(g, st, p, (uint32_t)d64);
  }
}

// InferenceResult::update_weight, src/inference_result.h:66-85, for one SGD step with
// stepsize step*t (t = truthiness under --noise_aware, else 1).
// Sequential schedule: the reference's read-modify-write, bit for bit.
// Parallel schedules: weights are frozen for the launch; the step's gradient AND its decay,
//   t*g + t*reg/(1+reg*step*t) * w      (l2)      t*g - reg*(w<0)/step      (l1, sic)
// are summed into grad_acc[wid]; k_apply_weights (dwg_sell.cuh) integrates the launch.
__device__ __forceinline__ void dev_update_weight(const DevState &st, uint32_t wid, const SweepParams &p,
                                                  double t, double grad) {
  double *w = st.weights;
  if (p.sequential) {
    const double step = __dmul_rn(p.stepsize, t);
    const double diff = __dmul_rn(step, grad);
    double x = w[wid];
    if (p.reg_l1) x = __dadd_rn(x, __dmul_rn(p.reg_param, x < 0 ? 1.0 : 0.0));
    else x = __dmul_rn(x, __ddiv_rn(1.0, __dadd_rn(1.0, __dmul_rn(p.reg_param, step))));
    w[wid] = __dsub_rn(x, diff);
    return;
  }
  double sum = t * grad;
  if (p.reg_param != 0.0) {
    const double x = w[wid];
    if (p.reg_l1) sum -= p.reg_param * (x < 0 ? 1.0 : 0.0) / p.stepsize;
    else sum += t * p.reg_param / (1.0 + p.reg_param * p.stepsize * t) * x;
  }
  if (sum != 0.0) atomicAdd(st.grad_acc + wid, sum);
}

// sgd_on_factor over one row, src/factor_graph.cc:243-260: for every non-fixed factor,
// gradient = potential(free chain) - potential(evid chain with this variable := x_evid)
template <typename AssignT>
__device__ __forceinline__ void dev_sgd_row(const DevGraph &g, const DevState &st, const SweepParams &p,
                                            uint32_t d, uint32_t row, bool cat, uint32_t key,
                                            uint32_t x_evid, uint32_t x_free, double t) {
  const AssignT *af = static_cast<const AssignT *>(st.assign_free);
  const AssignT *ae = static_cast<const AssignT *>(st.assign_evid);
  const uint32_t e0 = __ldg(g.row_ptr + row), e1 = __ldg(g.row_ptr + row + 1);
  EdgeCursor<AssignT> cur(g, d, g.row_oth ? __ldg(g.row_oth + row) : 0u);
  for (uint32_t e = e0; e < e1; ++e) {
    const uint8_t tag = __ldg(g.e_tag + e);
    const uint32_t wid = __ldg(g.e_wid + e);
    if (!__ldg(g.w_fixed + wid)) {  // :247
      const double fv = dev_fv(g, e);
      double se, sf, unused;
      cur.signs(tag, cat, key, ae, x_evid, x_evid, se, unused);
      cur.signs(tag, cat, key, af, x_free, x_free, sf, unused);
      const double grad = __dsub_rn(__dmul_rn(sf, fv), __dmul_rn(se, fv));  // :255-258
      dev_update_weight(st, wid, p, t, grad);
    }
    cur.advance(tag);
  }
}

// sample_sgd_single_variable (src/gibbs_sampler.h:127-149) for the variable at schedule position d
template <typename AssignT>
__device__ __forceinline__ void sweep_learn_var(const DevGraph &g, const DevState &st, const SweepParams &p, uint32_t d) {
  AssignT *af = static_cast<AssignT *>(st.assign_free);
  AssignT *ae = static_cast<AssignT *>(st.assign_evid);
  const AssignT *evid_val = static_cast<const AssignT *>(g.evid_val);
  const uint8_t fl = __ldg(g.var_flags + d);
  if (fl & VF_PAD) return;
  const bool cat = fl & VF_CATEGORICAL, evid = fl & VF_EVID, truthy = fl & VF_TRUTHINESS;
  uint32_t row0 = d, K = 1;
  if (g.var_row) {
    row0 = __ldg(g.var_row + d);
    K = __ldg(g.var_row + d + 1) - row0;
  }
  double r0, r1;
  dev_uniforms(p, d, r0, r1);
  // the free chain samples EVERY variable, evidence included (gibbs_sampler.h:137-139)
  const uint32_t pf = dev_draw_sample<AssignT, true>(g, d, cat, row0, K, af, st.weights, r0);
  af[d] = (AssignT)pf;
  // sample_evid, gibbs_sampler.h:171-190
  const uint32_t ev = (uint32_t)evid_val[d];
  uint32_t pe;
  if (!p.noise_aware && evid) {
    pe = ev;
  } else if (p.noise_aware && truthy) {
    double sum = 0;
    pe = 0;
    const uint32_t card = cat ? K : 2u;
    for (uint32_t i = 0; i < card; ++i) {
      sum = __dadd_rn(sum, g.truthiness[row0 + i]);
      if (sum >= r1) { pe = i; break; }
    }
  } else {
    pe = dev_draw_sample<AssignT, true>(g, d, cat, row0, K, ae, st.weights, r1);
  }
  ae[d] = (AssignT)pe;
  if (!p.learn_non_evidence && ((!p.noise_aware && !evid) || (p.noise_aware && !truthy)))
    return;  // :144-146
  // sgd_on_variable, src/factor_graph.cc:262-314
  if (!cat) {
    // Boolean: evidence_value is Variable::assignment_dense even for a query variable (:272)
    dev_sgd_row<AssignT>(g, st, p, d, row0, false, 0u, ev, pf, 1.0);
  } else {
    for (uint32_t val = 0; val < K; ++val) {
      if (!p.noise_aware && val != ev) continue;  // :280
      double t = 1.0;
      if (p.noise_aware) {
        t = g.truthiness ? g.truthiness[row0 + val] : 0.0;
        if (t <= 0.000001 && t >= -0.000001) continue;  // is_linear_zero :283
      }
      dev_sgd_row<AssignT>(g, st, p, d, row0 + val, true, val, val, pf, t);  // :288-292
      if (val == pf) continue;  // :304
      dev_sgd_row<AssignT>(g, st, p, d, row0 + pf, true, pf, val, pf, t);  // :306-311
    }
  }
}

template <typename AssignT>
__global__ void __launch_bounds__(SWEEP_THREADS)
k_sweep_learn(const DevGraph g, const DevState st, const SweepParams p) {
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;; tid += nthreads) {
    const uint64_t d64 = sweep_unit_var(p, tid);
    if (d64 >= p.d_end) {  // past the class: done once the whole 32-variable unit is
      if (d64 - (tid & 31u) >= p.d_end) break;
      continue;
    }
    sweep_learn_var<AssignT>(g, st, p, (uint32_t)d64);
  }
}

// The sequential schedule (tiny graphs: one chain walking the variables in id order, the
// reference's `-t 1 -c 1` semantics bit for bit) for n epochs in ONE launch: such graphs are pure
// launch latency otherwise (biased_coin: 18 variables x 4000 epochs).  The sweep counter of the
// Philox stream and the step size advance per epoch exactly as the host loop would advance them
// (stepsize *= decay, src/dimmwitted.cc:199-204).
template <typename AssignT>
__global__ void k_seq_epochs(const DevGraph g, const DevState st, SweepParams p, uint32_t n_epochs, int learn, double decay) {
  if (blockIdx.x || threadIdx.x) return;
  uint64_t sweep = ((uint64_t)p.sweep_hi << 32) | p.sweep_lo;
  for (uint32_t e = 0; e < n_epochs; ++e) {
    p.sweep_lo = (uint32_t)sweep; p.sweep_hi = (uint32_t)(sweep >> 32);
    for (uint32_t d = 0; d < g.nvar; ++d) {
      if (learn) sweep_learn_var<AssignT>(g, st, p, d);
      else sweep_infer_var<AssignT>(g, st, p, d);
    }
    ++sweep;
    if (learn) p.stepsize = __dmul_rn(p.stepsize, decay);
  }
}

// ---- parity hooks -----------------------------------------------------------------

// FactorGraph::potential (src/factor_graph.h:127-145) for explicit (d, proposal) queries
template <typename AssignT>
__global__ void k_eval_potentials(const DevGraph g, const DevState st, uint32_t n, const uint32_t *ds,
                                  const uint32_t *props, int evid_chain, double *out) {
  const AssignT *a = static_cast<const AssignT *>(evid_chain ? st.assign_evid : st.assign_free);
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t d = ds[i], prop = props[i];
  const bool cat = g.var_flags[d] & VF_CATEGORICAL;
  const uint32_t row0 = g.var_row ? g.var_row[d] : d;
  if (!cat) {
    double pos, neg;
    // a Boolean row answers proposals 0 and 1; any other proposal satisfies nothing
    const uint32_t e0 = g.row_ptr[row0], e1 = g.row_ptr[row0 + 1];
    EdgeCursor<AssignT> cur(g, d, g.row_oth ? g.row_oth[row0] : 0u);
    double pot = 0.0;
    for (uint32_t e = e0; e < e1; ++e) {
      const uint8_t tag = g.e_tag[e];
      double s0, s1;
      cur.signs(tag, false, 0u, a, prop, prop, s0, s1);
      cur.advance(tag);
      pot = dev_acc(pot, st.weights[g.e_wid[e]], s0, dev_fv(g, e));
    }
    (void)pos; (void)neg;
    out[i] = pot;
  } else {
    out[i] = dev_row_potential<AssignT, false>(g, d, row0 + prop, prop, a, st.weights);
  }
}

__global__ void k_draw(uint32_t n, const double *pots, uint32_t ld, const uint32_t *card,
                       const double *r, uint32_t *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *p = pots + (size_t)i * ld;
  out[i] = card[i] == 0 ? dev_draw_boolean(p[0], p[1], r[i]) : dev_draw_categorical_pots(p, card[i], r[i]);
}

__global__ void k_logadd(uint32_t n, const double *a, const double *b, double *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dev_logadd(a[i], b[i]);
}

// ---- replica averaging epilogue ---------------------------------------------------
// After the cross-replica SUM: w /= n_replicas (average_weights, src/inference_result.cc:75-79),
// lmax = max |w - prev|, l2sq = sum (w - prev)^2 (src/dimmwitted.cc:220-227), prev = w.
__global__ void k_finish_average(double *w, double *prev, uint32_t n, double inv_rep, int scale,
                                 double *norms /* [0]=lmax, [1]=l2sq */) {
  __shared__ double s_max[32], s_sum[32];
  double mx = 0.0, sq = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double x = w[i];
    if (scale) { x *= inv_rep; w[i] = x; }
    double diff = fabs(x - prev[i]);
    prev[i] = x;
    mx = fmax(mx, diff);
    sq += diff * diff;
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (lane == 0) { s_max[wp] = mx; s_sum[wp] = sq; }
  __syncthreads();
  if (wp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    mx = lane < nw ? s_max[lane] : 0.0;
    sq = lane < nw ? s_sum[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) {
      // non-negative doubles order like their bit patterns
      atomicMax(reinterpret_cast<unsigned long long *>(norms), (unsigned long long)__double_as_longlong(mx));
      atomicAdd(norms + 1, sq);
    }
  }
}

// dst[i] = src[idx[i]]: tallies from schedule order into ABI order (dwg_get_tallies)
__global__ void k_gather_u32(uint32_t *dst, const uint32_t *src, const uint32_t *idx, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[__ldg(idx + i)];
}

// narrow counters for the cross-replica exchange: FIELDS tallies of 32 / FIELDS bits per word
template <int FIELDS>
__global__ void k_pack_tallies(uint32_t *words, const uint32_t *t, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if ((uint64_t)i * FIELDS >= n) return;
  uint32_t w = 0;
#pragma unroll
  for (int f = 0; f < FIELDS; ++f) {
    const uint32_t k = i * FIELDS + f;
    if (k < n) w |= t[k] << (f * (32 / FIELDS));
  }
  words[i] = w;
}
template <int FIELDS>
__global__ void k_unpack_tallies(uint32_t *t, const uint32_t *words, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if ((uint64_t)i * FIELDS >= n) return;
  const uint32_t w = words[i];
#pragma unroll
  for (int f = 0; f < FIELDS; ++f) {
    const uint32_t k = i * FIELDS + f;
    if (k < n) t[k] = FIELDS == 1 ? w : (w >> (f * (32 / FIELDS))) & ((1u << (32 / FIELDS)) - 1u);
  }
}

// agg_nsamples of every variable in ABI order (src/gibbs_sampler.h:157,163: one count per sweep for
// every sampled variable): `sweeps`, or 0 for evidence when evidence is not sampled (evid = nullptr:
// every variable is sampled).  Written on the device so that a page-locked destination receives it
// by DMA: the copy engine fills host memory faster than the host's threads do.
__global__ void k_fill_nsamples(unsigned long long *out, const uint8_t *evid, unsigned long long sweeps, uint64_t n) {
  for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < n; v += (uint64_t)gridDim.x * blockDim.x)
    out[v] = (evid && __ldg(evid + v)) ? 0ull : sweeps;
}

// tallies (schedule order, u32) += host-provided counts (aggregate_marginals_from)
__global__ void k_add_u32(uint32_t *dst, const uint32_t *src, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

}  // namespace dwg
#endif
