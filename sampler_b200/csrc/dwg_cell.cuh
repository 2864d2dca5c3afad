// dwg_cell.cuh -- the production inference sweep for CATEGORICAL graphs (sm_100a): one warp per
// 32-variable slice of the categorical sliced-ELL blob (graph_image.h CELL_*), lane = variable.
//
//   k_cell_infer_staged  GibbsSamplerThread::sample -> sample_single_variable -> draw_sample,
//                        categorical branch (src/gibbs_sampler.cc:65-69, src/gibbs_sampler.h:151-169,
//                        217-246) with FactorGraph::potential per value (src/factor_graph.h:127-145)
//
// Memory behaviour: like the Boolean staged kernel (dwg_sell.cuh) a slice is ONE contiguous run of
// words, moved by one 1-D bulk-async copy (cp.async.bulk + mbarrier) into a per-warp ring in
// shared memory one slice ahead, and read once.  Per slice:
//   1. the chain values of the other variables of every two-variable factor are gathered in one
//      batch of independent L2 loads (u8 each) and kept as a bit mask per lane;
//   2. the lane walks its record column: a record carries the weight id and the factor's sign
//      table over (this variable satisfied, other variable satisfied); records of value k follow
//      those of value k - 1, a flag marks the first record of a row.  The row sums -- the
//      potentials of the K proposals, accumulated in the reference's factor order -- go to a
//      per-warp shared-memory array (lane-strided: no bank conflicts);
//   3. the draw.  The reference folds the potentials with logadd and walks r -= exp(p_k - sum)
//      (2 K fp64 transcendentals per variable: that alone would cost more than the memory
//      traffic).  An fp32 pass finds the value and how close r came to a cumulative boundary;
//      only when it came within 1e-4 (about one draw in 500 at K = 10) the exact fp64 scan decides.
//      Both passes see the same potentials, so the result is the exact scan's, bit for bit.
// Weights with id < n_tab sit in shared memory as three planes -w | 0 | +w (the record's sign code
// picks the plane).  No tensor cores: O(1) flop per 4 B of edge data.
#ifndef DWG_CELL_CUH_
#define DWG_CELL_CUH_
#include "dwg_sell.cuh"

namespace dwg {

struct CellGraph {
  uint32_t slice_begin, slice_end;
  const uint32_t *sl_off, *blob;
  uint32_t nweight, n_tab, kmax;
};

// shared-memory plan: signed weight table | mbarriers | rings | potentials (kmax x 32 doubles per warp)
__host__ __device__ inline uint32_t cell_buf_off(uint32_t n_tab, uint32_t stages, uint32_t stage_words, uint32_t nwarp) {
  return (sell_staged_smem(n_tab, stages, stage_words, nwarp) + 127u) & ~127u;
}
__host__ __device__ inline uint32_t cell_smem(uint32_t n_tab, uint32_t stages, uint32_t stage_words, uint32_t nwarp,
                                              uint32_t kmax) {
  return cell_buf_off(n_tab, stages, stage_words, nwarp) + nwarp * kmax * 32u * 8u;
}

// categorical draw on the K potentials buf[k * 32] (src/gibbs_sampler.h:217-246), exact fp64 form
__device__ __noinline__ uint32_t cell_draw_exact(const double *buf, uint32_t K, double r) {
  double sum = -100000.0;
  for (uint32_t k = 0; k < K; ++k) sum = dev_logadd(sum, buf[k * 32u]);
  for (uint32_t k = 0; k < K; ++k) {
    r = __dsub_rn(r, exp(__dsub_rn(buf[k * 32u], sum)));
    if (r <= 0) return k;
  }
  return K ? K - 1 : 0;  // the reference asserts here (:244)
}

// The same decision from an fp32 pass whenever it is safe: the cumulative probabilities c_k are
// computed to ~1e-5 (fp32 exp and sums; the reference's own sum is within 3e-7 of the true
// normaliser: its logadd drops terms below e^-18.42), so unless r lies within 1e-4 of a boundary
// c_k the exact scan stops at the same k.
__device__ __forceinline__ uint32_t cell_draw(const double *buf, uint32_t K, double r) {
  if (K <= 1) return 0;
  double md = buf[0];
  for (uint32_t k = 1; k < K; ++k) md = fmax(md, buf[k * 32u]);
  if (!(md > -90000.0 && md < 1e30)) return cell_draw_exact(buf, K, r);
  float S = 0.f;
  for (uint32_t k = 0; k < K; ++k) S += __expf((float)(buf[k * 32u] - md));
  const float target = (float)r * S, margin = 1e-4f * S;
  float c = 0.f;
  uint32_t pick = K;
  bool near = false;
  for (uint32_t k = 0; k < K; ++k) {
    c += __expf((float)(buf[k * 32u] - md));
    if (pick == K) {
      near |= fabsf(target - c) <= margin;
      if (target <= c) pick = k;
    }
  }
  if (near || pick == K) return cell_draw_exact(buf, K, r);
  return pick;
}

// sign * weight for a record's code (0 -> -w, 1 -> 0, 2 -> +w)
template <bool ALLHOT>
__device__ __forceinline__ double cell_signed_weight(uint32_t n_tab, const unsigned char *tab_s, uint32_t plane,
                                                     const double *w, uint32_t code, uint32_t wid) {
  if (ALLHOT) return *reinterpret_cast<const double *>(tab_s + (code * plane + 8u * wid));
  const bool hot = wid < n_tab;  // branch-free: one generic load from the table or from L2
  const double *ptr = hot ? reinterpret_cast<const double *>(tab_s + (code * plane + 8u * wid)) : w + wid;
  const double v = *ptr;
  return hot ? v : sell_term(code, v);
}

// potentials of the lane's K values into buf (lane-strided), records read from `src` (the staged
// copy of the slice, or the blob itself for a slice longer than a stage)
template <bool ALLHOT>
__device__ __forceinline__ void cell_potentials(const uint32_t *src, uint32_t n_rec, uint32_t n_oc, uint32_t lane,
                                                const uint8_t *a, uint32_t n_tab, const unsigned char *tab_s,
                                                uint32_t plane, const double *w, double *buf, uint32_t kmax) {
  const uint32_t *recs = src + SH_WORDS + 64u + lane, *oth = recs + 32u * n_rec;
  // 1. other variables' chain values: independent L2 gathers, four at a time.  (Requesting them a
  //    slice ahead, as the Boolean kernel does, costs 10 registers and measured 4.7 % SLOWER on C3.)
  uint32_t satmask = 0;
  for (uint32_t c0 = 0; c0 < n_oc; c0 += 4) {
    uint32_t ow[4], av[4];
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u) ow[u] = c0 + u < n_oc ? oth[32u * (c0 + u)] : 0u;
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u) av[u] = a[ow[u] >> 8];
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u)
      if (c0 + u < n_oc) satmask |= (av[u] == (ow[u] & 0xffu) ? 1u : 0u) << (c0 + u);
  }
  // 2. the record column: rows in value order, a row's records in factor-id order.  Branch-free: the
  //    running sum of the current row is stored after EVERY record (lane-strided, conflict-free), the
  //    last store of a row is its potential; a flagged record opens the next row with a fresh sum.
  //    (A branch per record on the row flag diverges in every warp: lanes change rows at different j.)
  double cur = 0.0;
  uint32_t k = 0xffffffffu, oc = 0;
#pragma unroll 4
  for (uint32_t j = 0; j < n_rec; ++j) {
    const uint32_t rec = recs[32u * j];
    const bool first = rec & CELL_ROW_BIT;  // (the first record of a lane always carries it: k becomes 0)
    k += first ? 1u : 0u;
    const uint32_t multi = (rec >> 30) & 1u, os = multi & (satmask >> oc);
    oc += multi;
    const uint32_t code = (rec >> (22u + 2u * (2u + os))) & 3u;  // this variable takes the row's value: self_sat = 1
    cur = __dadd_rn(first ? 0.0 : cur, cell_signed_weight<ALLHOT>(n_tab, tab_s, plane, w, code, rec & CELL_WID_MASK));
    buf[min(k, kmax - 1u) * 32u] = cur;  // (a padding lane has no flagged record: k stays -1, its stores land in the last row)
  }
}

template <bool ALLHOT>
__device__ __noinline__ void cell_unstaged_potentials(const CellGraph &g, uint32_t off, uint32_t n_rec, uint32_t n_oc,
                                                      uint32_t lane, const uint8_t *a, const unsigned char *tab_s,
                                                      uint32_t plane, const double *w, double *buf) {
  cell_potentials<ALLHOT>(g.blob + (size_t)off * 8, n_rec, n_oc, lane, a, g.n_tab, tab_s, plane, w, buf, g.kmax);
}

template <bool ALLHOT, int NT>
__global__ void __launch_bounds__(NT, 2)
k_cell_infer_staged(const CellGraph g, const SellState st, const SweepParams p, const uint32_t stages,
                    const uint32_t stage_words) {
  extern __shared__ __align__(128) unsigned char sell_smem[];
  // programmatic dependent launch: the prologue (weight table, mbarriers, offsets, first copies --
  // nothing a sweep writes) overlaps the tail of the previous colour class
  asm volatile("griddepcontrol.launch_dependents;");
  const unsigned char *tab_s = sell_smem;
  const uint32_t plane = sell_tab_plane(g.n_tab);
  const uint32_t *sm = reinterpret_cast<const uint32_t *>(sell_smem);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t bar0 = sell_smem_addr(sell_smem) + sell_staged_bar_off(g.n_tab) + warp * stages * 8u;
  const uint32_t ringw = sell_staged_ring_off(g.n_tab, stages, NT / 32) / 4u + warp * stages * stage_words;
  const uint32_t ring0 = sell_smem_addr(sell_smem) + ringw * 4u;
  double *buf = reinterpret_cast<double *>(sell_smem + cell_buf_off(g.n_tab, stages, stage_words, NT / 32)) +
                (size_t)warp * g.kmax * 32u + lane;
  if (lane == 0)
    for (uint32_t i = 0; i < stages; ++i) sell_mbar_init(bar0 + 8u * i, 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (uint32_t i = threadIdx.x; i < g.n_tab; i += blockDim.x) {
    const double w = __ldg(st.weights + i);
    *reinterpret_cast<double *>(sell_smem + 8u * i) = -w;
    *reinterpret_cast<double *>(sell_smem + plane + 8u * i) = 0.0;
    *reinterpret_cast<double *>(sell_smem + 2u * plane + 8u * i) = w;
  }
  __syncthreads();

  const uint64_t step = (uint64_t)((gridDim.x * blockDim.x) >> 5);
  const uint64_t s0 = g.slice_begin + (uint64_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (s0 >= g.slice_end) return;
  const uint32_t n_my = (uint32_t)((g.slice_end - 1 - s0) / step) + 1u;
  const uint64_t pol = sell_policy_evict_first();
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
  uint32_t stage = 0, parity = 0;
  // from here on the kernel reads and writes the chain and the tallies: the previous class must be done
  asm volatile("griddepcontrol.wait;" ::: "memory");
  uint64_t s64 = s0;
  for (uint32_t k = 0; k < n_my; ++k, s64 += step) {
    const uint32_t d = (uint32_t)s64 * 32u + lane;
    double r0, r1;
    dev_uniforms(p, d, r0, r1);
    sell_mbar_wait(bar0 + 8u * stage, parity);
    const uint32_t b = ringw + stage * stage_words;
    const uint4 h0 = *reinterpret_cast<const uint4 *>(sm + b);      // n_rec, n_oc, evid, pad
    const uint4 h1 = *reinterpret_cast<const uint4 *>(sm + b + 4);  // kmax, -, -, own offset
    const uint32_t n_rec = h0.x, n_oc = h0.y;
    // gibbs_sampler.h:157 -- a slice with nothing to sample is skipped
    const uint32_t active = ~h0.w & (p.sample_evidence ? 0xffffffffu : ~h0.z);
    if (active) {
      const uint32_t meta = sm[b + SH_WORDS + lane], row0 = sm[b + SH_WORDS + 32u + lane];  // header + meta + row0 always fit a stage
      if (SH_WORDS + 32u * (2u + n_rec + n_oc) > stage_words)
        cell_unstaged_potentials<ALLHOT>(g, h1.w, n_rec, n_oc, lane, st.assign_evid, tab_s, plane, st.weights, buf);
      else
        cell_potentials<ALLHOT>(sm + b, n_rec, n_oc, lane, st.assign_evid, g.n_tab, tab_s, plane, st.weights, buf, g.kmax);
      if ((active >> lane) & 1u) {
        const uint32_t prop = cell_draw(buf, meta & 0xffu, r0);
        st.assign_evid[d] = (uint8_t)prop;              // gibbs_sampler.h:160
        atomicAdd(st.tallies + row0 + prop, 1u);        // :163-167 (agg_nsamples is implicit)
      }
    }
    __syncwarp();  // every lane is done with this stage (and the potentials array) before the next copy lands
    if (k + stages < n_my) issue(k + stages, stage);
    if (++stage == stages) { stage = 0; parity ^= 1u; }
  }
}

// ---- learning sweep on the categorical sliced-ELL blob ---------------------------------------
// sample_sgd_single_variable for categorical variables (src/gibbs_sampler.h:127-149) with
// sgd_on_variable's categorical branch (src/factor_graph.cc:278-312, hard evidence:
// --noise_aware graphs keep the CSR kernels): the free chain draws every variable; the evid chain
// clamps evidence and draws the rest; a learning variable with evidence value e and free value f
// steps the weights of the factors in row e and, when f != e, in row f:
//     gradient = sign(free chain, this variable = f) - sign(evid chain, this variable := e)
// A record of row k has self_sat = (value == k) on either side, so both signs come out of its
// table.  The schedule is k_sell_learn_live's: live weights, a shared-memory copy
// refreshed every `period` slices, integer gradient accumulators (feature values are 1), steps of
// over-shared weights scaled, slices visited in ascending id block.
struct CellLearnExtra {
  const uint8_t *w_fixed;  // [nweight] internal order
  uint32_t kmax;
  int learn_non_evidence;
};

__host__ __device__ inline uint32_t cell_learn_buf_off(uint32_t n_hot, uint32_t stage_words) {
  return (uint32_t)((sell_learn_smem(n_hot, stage_words) + 127u) & ~(size_t)127u);
}
__host__ __device__ inline size_t cell_learn_smem(uint32_t n_hot, uint32_t stage_words, uint32_t kmax) {
  return (size_t)cell_learn_buf_off(n_hot, stage_words) + (size_t)(SELL_LEARN_THREADS / 32) * kmax * 32u * 8u;
}

// potentials of the lane's values on the chain whose others are summarised in satmask
__device__ __forceinline__ void cell_potentials_masked(const uint32_t *recs, uint32_t n_rec, uint32_t satmask,
                                                       const SellGraph &g, const double *w_s, const double *w,
                                                       double *buf, uint32_t kmax) {
  // (branch-free like cell_potentials: the running row sum is stored after every record)
  double cur = 0.0;
  uint32_t k = 0xffffffffu, oc = 0;
#pragma unroll 4
  for (uint32_t j = 0; j < n_rec; ++j) {
    const uint32_t rec = recs[32u * j];
    const bool first = rec & CELL_ROW_BIT;
    k += first ? 1u : 0u;
    const uint32_t multi = (rec >> 30) & 1u, os = multi & (satmask >> oc);
    oc += multi;
    const uint32_t code = (rec >> (22u + 2u * (2u + os))) & 3u;
    const uint32_t wid = rec & CELL_WID_MASK;
    const double wt = wid < g.n_hot ? w_s[wid] : __ldcg(w + wid);
    cur = __dadd_rn(first ? 0.0 : cur, sell_term(code, wt));
    buf[min(k, kmax - 1u) * 32u] = cur;
  }
}

__global__ void __launch_bounds__(SELL_LEARN_THREADS, 2)
k_cell_learn_live(const SellGraph g, const SellState st, const SweepParams p, const SellLearn L, const CellLearnExtra X,
                  const uint32_t stage_words) {
  extern __shared__ __align__(128) unsigned char sell_smem[];
  constexpr uint32_t ST = SELL_LEARN_STAGES;
  double *w_s = reinterpret_cast<double *>(sell_smem), *acc_s = w_s + g.n_hot;
  float *as_s = reinterpret_cast<float *>(acc_s + g.n_hot), *lam_s = as_s + g.n_hot, *debt_s = lam_s + g.n_hot;
  const uint32_t *sm = reinterpret_cast<const uint32_t *>(sell_smem);
  __shared__ uint32_t cnt_s;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t npos = L.vend - L.vbegin, rounds = (npos + warps - 1) / warps;
  const uint32_t n_my = gw < npos ? (npos - 1 - gw) / warps + 1u : 0u;
  const uint32_t bar0 = sell_smem_addr(sell_smem) + sell_learn_bar_off(g.n_hot) + warp * ST * 8u;
  const uint32_t ringw = sell_learn_ring_off(g.n_hot) / 4u + warp * ST * stage_words;
  const uint32_t ring0 = sell_smem_addr(sell_smem) + ringw * 4u;
  double *buf = reinterpret_cast<double *>(sell_smem + cell_learn_buf_off(g.n_hot, stage_words)) +
                (size_t)warp * X.kmax * 32u + lane;
  if (lane == 0)
    for (uint32_t i = 0; i < ST; ++i) sell_mbar_init(bar0 + 8u * i, 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (uint32_t i = threadIdx.x; i < g.n_hot; i += blockDim.x) {
    const bool fixed = __ldg(X.w_fixed + i);
    const double sc = sell_step_scale(L, i);
    as_s[i] = fixed ? 0.f : (float)(L.step * sc);
    const double per_update = L.reg_l1 ? L.decay : log1p(L.decay * sc);
    lam_s[i] = fixed ? 0.f : (float)((double)__ldg(L.w_N + i) * L.inv_nslice * per_update);
    acc_s[i] = 0.0;
    debt_s[i] = 0.f;
  }
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
  auto issue = [&](uint32_t k, uint32_t stage) {
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
  __syncthreads();
  if (n_my) {
    fetch(0, cur_s, cur_lo, cur_hi);
    fetch(32, nxt_s, nxt_lo, nxt_hi);
    const uint32_t pre = n_my < ST ? n_my : ST;
    for (uint32_t k = 0; k < pre; ++k) issue(k, k);
  }
  uint32_t stage = 0, parity = 0, k = 0;
  uint32_t proc_s = cur_s;
  for (uint32_t r0 = 0; r0 < rounds; r0 = r0 + L.period < r0 ? rounds : r0 + L.period) {
    __syncthreads();
    sell_refresh_weights(w_s, L.weights, g.n_hot);
    if (threadIdx.x == 0) cnt_s = 0;
    __syncthreads();
    uint32_t mine = 0;
    const uint32_t r1 = r0 + L.period < r0 ? rounds : min(rounds, r0 + L.period);
    for (uint32_t r = r0; r < r1 && k < n_my; ++r, ++k) {
      if ((k & 31u) == 0 && k) proc_s = cur_s;
      const uint32_t s = __shfl_sync(0xffffffffu, proc_s, k & 31u);
      sell_mbar_wait(bar0 + 8u * stage, parity);
      const uint32_t *b = sm + ringw + stage * stage_words;
      const uint4 h0 = *reinterpret_cast<const uint4 *>(b);      // n_rec, n_oc, evid, pad
      const uint4 h1 = *reinterpret_cast<const uint4 *>(b + 4);  // kmax, -, -, own offset
      const uint32_t n_rec = h0.x, n_oc = h0.y;
      // the records: the staged copy, or the blob itself for a slice longer than a stage
      const bool fits = SH_WORDS + 32u * (2u + n_rec + n_oc) <= stage_words;
      const uint32_t *src = fits ? b : g.blob + (size_t)h1.w * 8;
      const uint32_t *recs = src + SH_WORDS + 64u + lane, *oth = recs + 32u * n_rec;
      const uint32_t meta = b[SH_WORDS + lane], K = meta & 0xffu, ev = (meta >> 8) & 0xffu;
      const bool pad = (h0.w >> lane) & 1u, evid = (h0.z >> lane) & 1u;
      const uint32_t d = s * 32u + lane;
      double u0, u1;
      dev_uniforms(p, d, u0, u1);
      // both chains' values of the other variables, one batch of L2 gathers
      uint32_t fmask = 0, emask = 0;
      for (uint32_t c0 = 0; c0 < n_oc; c0 += 4) {
        uint32_t ow[4], af[4], ae[4];
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) ow[u] = c0 + u < n_oc ? oth[32u * (c0 + u)] : 0u;
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) { af[u] = __ldcg(st.assign_free + (ow[u] >> 8)); ae[u] = __ldcg(st.assign_evid + (ow[u] >> 8)); }
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u)
          if (c0 + u < n_oc) {
            fmask |= (af[u] == (ow[u] & 0xffu) ? 1u : 0u) << (c0 + u);
            emask |= (ae[u] == (ow[u] & 0xffu) ? 1u : 0u) << (c0 + u);
          }
      }
      // free chain: every variable (gibbs_sampler.h:137-139)
      cell_potentials_masked(recs, n_rec, fmask, g, w_s, L.weights, buf, X.kmax);
      const uint32_t pf = pad ? 0u : cell_draw(buf, K, u0);
      // evid chain (sample_evid, :171-190): hard evidence is clamped, everything else drawn
      uint32_t pe = ev;
      if (__any_sync(0xffffffffu, !pad && !evid)) {
        cell_potentials_masked(recs, n_rec, emask, g, w_s, L.weights, buf, X.kmax);
        if (!pad && !evid) pe = cell_draw(buf, K, u1);
      }
      if (!pad) { st.assign_free[d] = (uint8_t)pf; st.assign_evid[d] = (uint8_t)pe; }
      // sgd_on_variable, categorical (factor_graph.cc:278-312): rows ev and (if different) pf
      const bool learn = !pad && (X.learn_non_evidence || evid);
      if (__any_sync(0xffffffffu, learn)) {
        uint32_t row = 0xffffffffu, oc = 0;
        for (uint32_t j = 0; j < n_rec; ++j) {
          const uint32_t rec = recs[32u * j];
          row += (rec >> 31) & 1u;  // CELL_ROW_BIT: the first record of the next value
          const uint32_t multi = (rec >> 30) & 1u;
          const uint32_t osf = multi & (fmask >> oc), ose = multi & (emask >> oc);
          oc += multi;
          if (!learn || (row != ev && row != pf) || (rec & ~CELL_ROW_BIT) == CELL_NOP) continue;
          const uint32_t tab = rec >> 22, ssf = pf == row, sse = ev == row;
          const int cf = (int)((tab >> (2u * (2u * ssf + osf))) & 3u), ce = (int)((tab >> (2u * (2u * sse + ose))) & 3u);
          const uint32_t wid = rec & CELL_WID_MASK;
          if (wid < g.n_hot) {
            if (cf != ce) atomicAdd(reinterpret_cast<int *>(acc_s + wid), cf - ce);
          } else if (cf != ce && !__ldg(X.w_fixed + wid)) {  // beyond the copy: the step straight into the live vector (decay: k_cold_decay)
            atomicAdd(L.weights + wid, -L.step * sell_step_scale(L, wid) * (double)(cf - ce));
          }
        }
      }
      ++mine;
      __syncwarp();
      if (k + ST < n_my) issue(k + ST, stage);
      if (++stage == ST) { stage = 0; parity ^= 1u; }
    }
    if (lane == 0 && mine) atomicAdd(&cnt_s, mine);
    __syncthreads();
    const uint32_t cnt = cnt_s;
    const bool last = r1 >= rounds;
    for (uint32_t i = threadIdx.x; i < g.n_hot; i += blockDim.x) {
      const double a = sell_acc_take<FV_ONE>(acc_s, i), x = w_s[i];
      const float debt = debt_s[i] + lam_s[i] * (float)cnt;
      if (a == 0.0 && debt < 1e-3f && !last) { debt_s[i] = debt; continue; }
      debt_s[i] = 0.f;
      const double dec = L.reg_l1 ? (double)debt * (x < 0 ? 1.0 : 0.0) : x * expm1(-(double)debt);
      const double dlt = dec - (double)as_s[i] * a;
      if (dlt != 0.0) atomicAdd(L.weights + i, dlt);
    }
  }
}

}  // namespace dwg
#endif
