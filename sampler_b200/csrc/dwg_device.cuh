// dwg_device.cuh -- device-side math of the Gibbs hot path (sm_100a).
//
// Every function names the reference code it restates.  All potential / draw
// arithmetic is fp64 with explicit round-to-nearest intrinsics (__dmul_rn/__dadd_rn),
// so that nvcc cannot contract a*b+c into an FMA: the reference is baseline x86-64
// (mul and add round separately) and the draw must be bit-exact given the uniform.
// No tensor cores anywhere: the work is a sparse gather with O(1) flops per edge.
#ifndef DWG_DEVICE_CUH_
#define DWG_DEVICE_CUH_
#include <cstdint>
#include <cuda_runtime.h>

#include "graph_image.h"

namespace dwg {

// ---------------------------------------------------------------------------
// device views
// ---------------------------------------------------------------------------
struct DevGraph {
  uint32_t nvar, nrow;
  const uint8_t *var_flags;   // [nvar]
  const void *evid_val;       // [nvar] AssignT
  const uint32_t *var_row;    // [nvar+1] or null (Boolean-only graph: row == d)
  const uint32_t *row_ptr;    // [nrow+1]
  const uint32_t *row_oth;    // [nrow+1] or null (unary-only graph)
  const double *truthiness;   // [nrow] or null
  const uint32_t *e_wid;
  const float *e_fv32;        // one of these two, or neither (all feature values 1.0)
  const double *e_fv64;
  const uint8_t *e_tag;
  const uint32_t *oth;
  int oth_packed;
  const uint8_t *w_fixed;     // [nweight]
};

struct DevState {
  void *assign_free, *assign_evid;  // AssignT[nvar]
  double *weights;                  // f64[nweight], internal weight order
  double *grad_acc;                 // f64[nweight] gradient sums of the current learning launch
  uint32_t *tallies;                // u32[ndrow]
};

struct SweepParams {
  uint32_t d_begin, d_end;  // variables of this launch (one colour class, schedule order)
  uint32_t key0, key1;      // Philox key: seed ^ replica stream
  uint32_t sweep_lo, sweep_hi;  // sweep counter (part of the Philox counter)
  int sample_evidence, learn_non_evidence, noise_aware, reg_l1, sequential;
  double reg_param, stepsize;
  const double *r_free, *r_evid;  // explicit uniforms in schedule order (parity hooks) or null
  // learning chunks: this launch sweeps the 32-variable units  phase, phase + stride, ... of the class
  uint32_t unit_stride, unit_phase;
};

// variable handled by global thread `tid` in round `k` of a (possibly strided) class sweep
__host__ __device__ inline uint64_t sweep_unit_var(const SweepParams &p, uint64_t tid) {
  return (uint64_t)p.d_begin + (((tid >> 5) * p.unit_stride + p.unit_phase) << 5) + (tid & 31u);
}

// ---------------------------------------------------------------------------
// counter-based RNG: Philox4x32-10 (Salmon et al., SC'11).  One call per variable
// per sweep: counter = (d, purpose, sweep_lo, sweep_hi), key = (seed, replica).
// Replaces the per-thread erand48 stream (src/gibbs_sampler.cc:49, gibbs_sampler.h:203).
// ---------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0,1) (erand48 gives 48 bits; same support, finer grid)
__host__ __device__ inline double u01(uint32_t hi, uint32_t lo) {
  uint64_t x = ((uint64_t)hi << 32 | lo) >> 11;
  return (double)x * (1.0 / 9007199254740992.0);
}

#ifdef __CUDACC__

__device__ __forceinline__ void dev_uniforms(const SweepParams &p, uint32_t d, double &r0, double &r1) {
  if (p.r_free) {
    r0 = p.r_free[d];
    r1 = p.r_evid ? p.r_evid[d] : 0.0;
    return;
  }
  uint32_t o[4];
  philox4x32_10(d, 0u, p.sweep_lo, p.sweep_hi, p.key0, p.key1, o);
  r0 = u01(o[0], o[1]);
  r1 = u01(o[2], o[3]);
}

// log2(1 + k) for k < 64 as the HOST libm rounds it (filled once by dwg_api.cu): FUNC_RATIO
// (src/factor.h:264-277) takes log2 of a small integer, and CUDA's log2 may differ from
// glibc's in the last bit.  Larger arguments use the device log2 (<= 1 ulp).
constexpr int DWG_LOG2_TABLE = 64;
__constant__ double c_log2_1p[DWG_LOG2_TABLE];

// ---------------------------------------------------------------------------
// logadd, src/common.h:118-132 (LOG_2 and the -18.42 cut-off as literally spelled there)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double dev_logadd(double log_a, double log_b) {
  if (log_a < log_b) {
    double t = log_a; log_a = log_b; log_b = t;
  } else if (log_a <= log_b && log_b <= log_a) {
    return __dadd_rn(0.693147180559945, log_a);
  }
  double nad = __dsub_rn(log_b, log_a);
  if (nad < -18.42) return log_a;
  return __dadd_rn(log_a, log1p(exp(nad)));
}

// Boolean draw, src/gibbs_sampler.h:198-215:  r * (1 + exp(neg - pos)) < 1
__device__ __forceinline__ uint32_t dev_draw_boolean(double pos, double neg, double r) {
  return __dmul_rn(r, __dadd_rn(1.0, exp(__dsub_rn(neg, pos)))) < 1.0 ? 1u : 0u;
}

// ---------------------------------------------------------------------------
// Factor::potential sign table, src/factor.h:112-299, from three counts:
//   n     arity,  nsb  satisfied positions among the body (all but the last),
//   head  the last position is satisfied.
// EQUAL ("all flags equal the first's") == all satisfied or none satisfied.
// LINEAR/RATIO/LOGICAL: sum over the body of ((!sat) || head) == head ? n-1 : n-1-nsb.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double dev_sign(uint32_t fc, uint32_t n, uint32_t nsb, bool head) {
  const uint32_t nbody = n - 1;
  const bool body_all = nsb == nbody;
  switch (fc) {
    case FC_AND: return (body_all && head) ? 1.0 : -1.0;
    case FC_AND_CATEGORICAL: return (body_all && head) ? 1.0 : 0.0;
    case FC_OR: return (nsb > 0 || head) ? 1.0 : -1.0;
    case FC_EQUAL: { uint32_t t = nsb + (head ? 1u : 0u); return (t == 0 || t == n) ? 1.0 : -1.0; }
    case FC_IMPLY_NATURAL: return body_all ? (head ? 1.0 : -1.0) : 0.0;
    case FC_IMPLY_MLN: return body_all ? (head ? 1.0 : 0.0) : 1.0;
    default: break;
  }
  // LINEAR / RATIO / LOGICAL
  if (n == 1) return head ? 1.0 : 0.0;  // log2(1+1)=1, log2(1+0)=0 exactly
  const double res = head ? (double)nbody : (double)(nbody - nsb);
  if (fc == FC_LINEAR) return res;
  if (fc == FC_RATIO) return res < (double)DWG_LOG2_TABLE ? c_log2_1p[(int)res] : log2(__dadd_rn(1.0, res));
  return res > 0.0 ? 1.0 : 0.0;  // FC_LOGICAL
}

template <typename AssignT>
struct EdgeCursor {
  const DevGraph &g;
  const uint32_t self;  // d of the variable being sampled
  uint32_t oc;          // cursor into the others stream (entries)
  __device__ EdgeCursor(const DevGraph &g_, uint32_t self_, uint32_t oc_) : g(g_), self(self_), oc(oc_) {}

  __device__ __forceinline__ void entry(uint32_t i, uint32_t &d, uint32_t &eq) const {
    if (g.oth_packed) {
      uint32_t w = __ldg(g.oth + i);
      d = w >> 1; eq = w & 1u;
    } else {
      uint2 w = __ldg(reinterpret_cast<const uint2 *>(g.oth) + i);
      d = w.x; eq = w.y;
    }
  }

  // number of others-stream entries behind one edge record
  __device__ __forceinline__ uint32_t span(uint8_t tag) const {
    const uint32_t form = tag >> TAG_FORM_SHIFT;
    if (form != FORM_GENERAL) return form;
    return 1u + (g.oth_packed ? __ldg(g.oth + oc) : __ldg(g.oth + 2 * (size_t)oc));
  }
  __device__ __forceinline__ void advance(uint8_t tag) { oc += span(tag); }

  // Signs of the factor behind one edge record when the sampled variable takes value
  // x0 (-> s0) and x1 (-> s1), every other variable read from chain `a`
  // (is_variable_satisfied, src/factor.h:94-100).  self_key: this row's dense value for a
  // categorical row.  Does not move the cursor.
  __device__ __forceinline__ void signs(uint8_t tag, bool cat, uint32_t self_key, const AssignT *a,
                                        uint32_t x0, uint32_t x1, double &s0, double &s1) const {
    const uint32_t fc = tag & TAG_FUNC_MASK;
    const uint32_t form = tag >> TAG_FORM_SHIFT;
    if (form != FORM_GENERAL) {
      const uint32_t self_eq = cat ? self_key : ((tag & TAG_SELF_EQ1) ? 1u : 0u);
      const bool self_head = tag & TAG_SELF_HEAD;
      uint32_t nsb = 0;
      bool head_oth = false;
      for (uint32_t j = 0; j < form; ++j) {
        uint32_t d, eq;
        entry(oc + j, d, eq);
        bool s = (uint32_t)a[d] == eq;
        if (!self_head && j == form - 1) head_oth = s;
        else nsb += s;
      }
      const bool m0 = x0 == self_eq, m1 = x1 == self_eq;
      if (self_head) {
        s0 = dev_sign(fc, form + 1, nsb, m0);
        s1 = dev_sign(fc, form + 1, nsb, m1);
      } else {
        s0 = dev_sign(fc, form + 1, nsb + m0, head_oth);
        s1 = dev_sign(fc, form + 1, nsb + m1, head_oth);
      }
      return;
    }
    // general form: [n][pos 0]...[pos n-1], self positions included
    const uint32_t n = span(tag) - 1u;
    uint32_t nsb0 = 0, nsb1 = 0;
    bool h0 = false, h1 = false;
    for (uint32_t i = 0; i < n; ++i) {
      uint32_t d, eq;
      entry(oc + 1 + i, d, eq);
      bool t0, t1;
      if (d == self) { t0 = eq == x0; t1 = eq == x1; }
      else { t0 = t1 = (uint32_t)a[d] == eq; }
      if (i == n - 1) { h0 = t0; h1 = t1; }
      else { nsb0 += t0; nsb1 += t1; }
    }
    s0 = dev_sign(fc, n, nsb0, h0);
    s1 = dev_sign(fc, n, nsb1, h1);
  }
};

__device__ __forceinline__ double dev_fv(const DevGraph &g, uint32_t e) {
  if (g.e_fv32) return (double)__ldg(g.e_fv32 + e);
  if (g.e_fv64) return __ldg(g.e_fv64 + e);
  return 1.0;
}

// pot += weight * (sign * feature_value), src/factor_graph.h:140-142 + src/factor.h:65-66
__device__ __forceinline__ double dev_acc(double pot, double w, double sign, double fv) {
  return __dadd_rn(pot, __dmul_rn(w, __dmul_rn(sign, fv)));
}

// FactorGraph::potential for a Boolean variable, both proposals in one pass over its row
// (src/factor_graph.h:127-145; draw_sample evaluates proposal 1 then 0, gibbs_sampler.h:201-202)
// LIVE: the launch itself rewrites weights (the sequential schedule's per-step update), so they
// are read with ordinary coherent loads; otherwise they are constant for the launch and go
// through the read-only path.
template <bool LIVE>
__device__ __forceinline__ double dev_weight(const double *w, uint32_t wid) {
  return LIVE ? w[wid] : __ldg(w + wid);
}

template <typename AssignT, bool LIVE>
__device__ __forceinline__ void dev_bool_potentials(const DevGraph &g, uint32_t d, uint32_t row,
                                                    const AssignT *a, const double *w,
                                                    double &pos, double &neg) {
  const uint32_t e0 = __ldg(g.row_ptr + row), e1 = __ldg(g.row_ptr + row + 1);
  EdgeCursor<AssignT> cur(g, d, g.row_oth ? __ldg(g.row_oth + row) : 0u);
  pos = 0.0; neg = 0.0;
  for (uint32_t e = e0; e < e1; ++e) {
    const uint8_t tag = __ldg(g.e_tag + e);
    const double wt = dev_weight<LIVE>(w, __ldg(g.e_wid + e));
    const double fv = dev_fv(g, e);
    double s0, s1;
    cur.signs(tag, false, 0u, a, 0u, 1u, s0, s1);
    cur.advance(tag);
    pos = dev_acc(pos, wt, s1, fv);
    neg = dev_acc(neg, wt, s0, fv);
  }
}

// FactorGraph::potential(var, k) for one categorical row
template <typename AssignT, bool LIVE>
__device__ __forceinline__ double dev_row_potential(const DevGraph &g, uint32_t d, uint32_t row,
                                                    uint32_t k, const AssignT *a, const double *w) {
  const uint32_t e0 = __ldg(g.row_ptr + row), e1 = __ldg(g.row_ptr + row + 1);
  EdgeCursor<AssignT> cur(g, d, g.row_oth ? __ldg(g.row_oth + row) : 0u);
  double pot = 0.0;
  for (uint32_t e = e0; e < e1; ++e) {
    const uint8_t tag = __ldg(g.e_tag + e);
    const double wt = dev_weight<LIVE>(w, __ldg(g.e_wid + e));
    const double fv = dev_fv(g, e);
    double s0, s1;
    cur.signs(tag, true, k, a, k, k, s0, s1);
    cur.advance(tag);
    pot = dev_acc(pot, wt, s0, fv);
  }
  return pot;
}

constexpr int DWG_CAT_BUF = 32;  // potentials kept in local memory up to this cardinality

// Categorical draw, src/gibbs_sampler.h:217-246: sum = sequential logadd from -100000,
// then the first i with (r -= exp(pot_i - sum)) <= 0.  Above DWG_CAT_BUF values the row
// potentials are recomputed in the second pass instead of buffered.  Where the reference
// asserts (no i found, :244) the last value is returned.
template <typename AssignT, bool LIVE>
__device__ uint32_t dev_draw_categorical(const DevGraph &g, uint32_t d, uint32_t row0, uint32_t K,
                                         const AssignT *a, const double *w, double r) {
  double buf[DWG_CAT_BUF];
  double sum = -100000.0;
  for (uint32_t k = 0; k < K; ++k) {
    double p = dev_row_potential<AssignT, LIVE>(g, d, row0 + k, k, a, w);
    if (k < DWG_CAT_BUF) buf[k] = p;
    sum = dev_logadd(sum, p);
  }
  for (uint32_t k = 0; k < K; ++k) {
    double p = k < DWG_CAT_BUF ? buf[k] : dev_row_potential<AssignT, LIVE>(g, d, row0 + k, k, a, w);
    r = __dsub_rn(r, exp(__dsub_rn(p, sum)));
    if (r <= 0) return k;
  }
  return K ? K - 1 : 0;
}

// the same decision on explicit potentials (parity hook dwg_draw)
__device__ inline uint32_t dev_draw_categorical_pots(const double *pots, uint32_t K, double r) {
  double sum = -100000.0;
  for (uint32_t k = 0; k < K; ++k) sum = dev_logadd(sum, pots[k]);
  for (uint32_t k = 0; k < K; ++k) {
    r = __dsub_rn(r, exp(__dsub_rn(pots[k], sum)));
    if (r <= 0) return k;
  }
  return 0xffffffffu;  // the reference's assert :244
}

#endif  // __CUDACC__
}  // namespace dwg
#endif
