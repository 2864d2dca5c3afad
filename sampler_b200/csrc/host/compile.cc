// compile.cc -- decoded records -> GraphImage (the device layout).
//
// Reproduces, quirks included, what the reference does between reading the files and
// the first sweep: load_variables / load_domains / load_factors value mapping
// (src/binary_format.cc:72-226), safety_check (src/factor_graph.cc:201-219) and
// construct_index (src/factor_graph.cc:90-199: per variable, entries sorted by
// (dense value, factor id), consecutive duplicates dropped).  On top of that it picks the
// sweep schedule (greedy colouring), permutes variables into schedule order and flattens
// factor -> vifs into per-row edge records (graph_image.h).
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <vector>

#include "../../../include/dwg.h"
#include "../graph_image.h"
#include "parallel.h"

namespace dwg {

RawView view_of(const RawGraph &r) {
  RawView v;
  v.nvar = r.var_role.size();
  v.ndom = r.dom_vid.size();
  v.nfac = r.fac_func.size();
  v.nweight = r.w_init.size();
  v.var_role = r.var_role.data(); v.var_init = r.var_init.data();
  v.var_dtype = r.var_dtype.data(); v.var_card = r.var_card.data();
  v.dom_vid = r.dom_vid.data(); v.dom_ptr = r.dom_ptr.data(); v.dom_val = r.dom_val.data();
  v.dom_truth = r.dom_truth.data();
  v.fac_func = r.fac_func.data(); v.fac_ptr = r.fac_ptr.data(); v.edge_vid = r.edge_vid.data();
  v.edge_eq = r.edge_eq.data(); v.fac_wid = r.fac_wid.data(); v.fac_fv = r.fac_fv.data();
  v.w_init = r.w_init.data(); v.w_fixed = r.w_fixed.data();
  return v;
}

namespace {

constexpr uint64_t INVALID = ~0ull;
constexpr uint64_t U32_LIMIT = 0xffffffffull;

int func_code(int32_t f) {
  switch (f) {
    case DWG_FUNC_IMPLY_NATURAL: return FC_IMPLY_NATURAL;
    case DWG_FUNC_OR: return FC_OR;
    case DWG_FUNC_AND:
    case DWG_FUNC_ISTRUE: return FC_AND;
    case DWG_FUNC_EQUAL: return FC_EQUAL;
    case DWG_FUNC_LINEAR: return FC_LINEAR;
    case DWG_FUNC_RATIO: return FC_RATIO;
    case DWG_FUNC_LOGICAL: return FC_LOGICAL;
    case DWG_FUNC_AND_CATEGORICAL: return FC_AND_CATEGORICAL;
    case DWG_FUNC_IMPLY_MLN: return FC_IMPLY_MLN;
    default: return -1;  // src/factor.h:81-84 aborts on these at the first sweep
  }
}

struct DomEntry {
  uint64_t val;
  uint32_t idx;
};

// sparse -> dense value maps (Variable::get_domain_index, src/variable.h:118): per domain
// block a (value, index) table sorted by value.  unordered_map insert-or-assign
// (src/binary_format.cc:215): a repeated value keeps its LAST index.
struct Domains {
  std::vector<int64_t> block_of;  // per var, -1 = identity domain
  std::vector<uint64_t> tab_ptr;  // per block
  std::vector<DomEntry> tab;
  uint64_t lookup(uint64_t v, uint64_t sparse) const {
    int64_t b = block_of[v];
    if (b < 0) return sparse;
    const DomEntry *lo = tab.data() + tab_ptr[b], *hi = tab.data() + tab_ptr[b + 1];
    const DomEntry *it = std::lower_bound(
        lo, hi, sparse, [](const DomEntry &e, uint64_t x) { return e.val < x; });
    if (it == hi || it->val != sparse) return INVALID;
    return it->idx;
  }
};

struct Inc {  // TempValueFactor, src/variable.h:14-29
  uint32_t key;  // dense value (0 for Boolean, src/binary_format.cc:166-169)
  uint32_t fid;
};

}  // namespace

// DWG_COMPILE_TIMING=1: wall time of each phase on stderr
struct PhaseTimer {
  bool on = getenv("DWG_COMPILE_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char *what) {
    if (!on) return;
    auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[compile] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

// The chromatic schedule: a proper colouring of the conflict graph "two variables share a factor"
// (greedy in id order), balanced so that every class can fill the GPU.  Returns the colour count;
// color[v] is written for every variable.  inc / deg / inc_end: the incidences of variable v are
// inc[deg[v] .. inc_end[v]) (construct_index order).
template <typename IncVec, typename DegVec, typename EndVec>
static uint32_t colour_variables(const RawView &raw, uint64_t nvar, const IncVec &inc, const DegVec &deg,
                                 const EndVec &inc_end, std::vector<uint32_t> &color) {
  uint32_t ncolor = 1;
  // conflict lists (the other variables of every incident factor; repeats are harmless), built
  // in parallel; the two greedy passes below then touch nothing but these and color[]
  std::vector<uint64_t> nb_ptr(nvar + 1, 0);
  auto for_neighbours = [&](uint64_t v, auto &&fn) {
    uint32_t last_f = 0xffffffffu;
    for (uint32_t i = deg[v]; i < inc_end[v]; ++i) {
      uint32_t f = inc[i].fid;
      if (f == last_f) continue;
      last_f = f;
      const uint64_t a = raw.fac_ptr[f], b = raw.fac_ptr[f + 1];
      if (b - a < 2) continue;
      for (uint64_t e = a; e < b; ++e)
        if (raw.edge_vid[e] != v) fn((uint32_t)raw.edge_vid[e]);
    }
  };
  parallel_for(nvar, [&](uint64_t v0, uint64_t v1) {
    for (uint64_t v = v0; v < v1; ++v) {
      uint64_t n = 0;
      for_neighbours(v, [&](uint32_t) { ++n; });
      nb_ptr[v + 1] = n;
    }
  });
  for (uint64_t v = 0; v < nvar; ++v) nb_ptr[v + 1] += nb_ptr[v];
  BulkVec<uint32_t> nb(nb_ptr[nvar]);
  parallel_for(nvar, [&](uint64_t v0, uint64_t v1) {
    for (uint64_t v = v0; v < v1; ++v) {
      uint64_t o = nb_ptr[v];
      for_neighbours(v, [&](uint32_t u) { nb[o++] = u; });
    }
  });
  // Pass 1, first fit, finds how many colours K the graph needs; pass 2 recolours with the
  // least-loaded admissible colour among those K, so that every class is large enough to
  // fill the GPU (first fit alone leaves a tail of tiny classes = latency-bound launches).
  // Pass 1 in parallel wavefronts.  The first-fit colour of v depends only on the colours of its
  // lower-numbered neighbours, so colouring v once all of those are coloured gives exactly the
  // sequential result, whatever the interleaving: every round colours the pending variables that
  // are ready (in id order inside each thread's range), the rest wait for the next round.  Graphs
  // with long id-ordered dependency chains (v coupled to v+1) make little progress per round:
  // the remainder is then finished sequentially.
  {
    std::fill(color.begin(), color.end(), 0xffffffffu);
    auto first_fit = [&](uint32_t v, std::vector<uint32_t> &stamp) -> bool {  // false: not ready yet
      for (uint64_t i = nb_ptr[v]; i < nb_ptr[v + 1]; ++i) {
        const uint32_t u = nb[i];
        const uint32_t c = __atomic_load_n(&color[u], __ATOMIC_ACQUIRE);
        if (c == 0xffffffffu) {
          if (u < v) return false;
          continue;
        }
        if (c >= stamp.size()) stamp.resize(c + 2, 0);
        stamp[c] = v + 1;
      }
      uint32_t c = 0;
      while (c < stamp.size() && stamp[c] == v + 1) ++c;
      __atomic_store_n(&color[v], c, __ATOMIC_RELEASE);
      return true;
    };
    const unsigned T = host_threads();
    std::vector<uint32_t> pending(nvar);
    std::iota(pending.begin(), pending.end(), 0u);
    while (!pending.empty()) {
      if (T == 1 || pending.size() < (1u << 14)) {
        std::vector<uint32_t> stamp(ncolor + 2, 0);
        for (uint32_t v : pending) first_fit(v, stamp);  // in id order every variable is ready
        break;
      }
      std::vector<std::vector<uint32_t>> later(T);
      parallel_ranges(pending.size(), T, [&](uint64_t i0, uint64_t i1, unsigned t) {
        std::vector<uint32_t> stamp(64, 0);
        for (uint64_t i = i0; i < i1; ++i)
          if (!first_fit(pending[i], stamp)) later[t].push_back(pending[i]);
      });
      std::vector<uint32_t> rest;
      for (auto &l : later) rest.insert(rest.end(), l.begin(), l.end());  // still ascending
      const bool stalled = rest.size() * 8 > pending.size() * 7;
      pending.swap(rest);
      if (stalled && !pending.empty()) {
        std::vector<uint32_t> stamp(64, 0);
        for (uint32_t v : pending) first_fit(v, stamp);
        break;
      }
    }
    uint32_t mx = 0;
    for (uint64_t v = 0; v < nvar; ++v) mx = std::max(mx, color[v]);
    ncolor = mx + 1;
  }
  // Pass 2 (sequential: the least-loaded choice depends on every earlier one)
  std::vector<uint32_t> stamp;  // stamp[c] == v+1 : colour c is taken by a neighbour of v
  std::vector<uint64_t> load;
  {
    const uint32_t K = ncolor;
    std::fill(color.begin(), color.end(), 0xffffffffu);
    stamp.assign(K + 1, 0);
    load.assign(K + 1, 0);
    ncolor = 1;
    for (uint64_t v = 0; v < nvar; ++v) {
      // the colours of the neighbours are random reads: ask for those of variable v + 16 now
      if (v + 16 < nvar)
        for (uint64_t i = nb_ptr[v + 16]; i < nb_ptr[v + 17]; ++i) __builtin_prefetch(&color[nb[i]]);
      for (uint64_t i = nb_ptr[v]; i < nb_ptr[v + 1]; ++i) {
        uint32_t c = color[nb[i]];
        if (c == 0xffffffffu) continue;
        stamp[c] = (uint32_t)v + 1;
      }
      uint32_t c = 0xffffffffu;
      for (uint32_t k = 0; k < K; ++k)
        if (stamp[k] != (uint32_t)v + 1 && (c == 0xffffffffu || load[k] < load[c])) c = k;
      if (c == 0xffffffffu) {  // every planned colour is taken around v: open a new one
        c = K;
        while (c < stamp.size() && stamp[c] == (uint32_t)v + 1) ++c;
      }
      if (c + 1 >= stamp.size()) { stamp.resize(c + 2, 0); load.resize(c + 2, 0); }
      color[v] = c;
      ++load[c];
      ncolor = std::max(ncolor, c + 1);
    }
    // Pass 3: the balancing pass strands a few variables whose neighbours hold all K planned colours in
    // an overflow class of their own (C2: 448 of 10 M) -- a whole launch boundary per sweep for a class
    // that cannot fill one SM.  Each is brought back by moving the neighbours that hold one planned
    // colour c to another colour free around THEM (a one-step Kempe move); v then takes c.
    if (ncolor > K) {
      uint32_t tick = (uint32_t)nvar + 1;  // stamps beyond the per-variable ones of pass 2
      std::vector<uint32_t> mark(stamp.size(), 0);
      auto next_tick = [&]() {
        if (++tick == 0) { std::fill(mark.begin(), mark.end(), 0u); tick = 1; }
        return tick;
      };
      // a planned colour != avoid that no neighbour of u holds (the least loaded one), or none
      auto free_colour = [&](uint32_t u, uint32_t avoid) -> uint32_t {
        const uint32_t t = next_tick();
        for (uint64_t i = nb_ptr[u]; i < nb_ptr[u + 1]; ++i) mark[color[nb[i]]] = t;
        uint32_t best = 0xffffffffu;
        for (uint32_t k = 0; k < K; ++k)
          if (k != avoid && mark[k] != t && (best == 0xffffffffu || load[k] < load[best])) best = k;
        return best;
      };
      std::vector<std::pair<uint32_t, uint32_t>> moves;
      for (uint64_t v = 0; v < nvar; ++v) {
        if (color[v] < K) continue;
        for (uint32_t c = 0; c < K; ++c) {
          moves.clear();
          bool ok = true;
          for (uint64_t i = nb_ptr[v]; i < nb_ptr[v + 1] && ok; ++i) {
            const uint32_t u = nb[i];
            if (color[u] != c) continue;
            bool seen = false;  // (a neighbour listed twice: two shared factors)
            for (auto &m : moves) seen |= m.first == u;
            if (seen) continue;
            const uint32_t c2 = free_colour(u, c);
            if (c2 == 0xffffffffu) ok = false;
            else moves.emplace_back(u, c2);
          }
          if (!ok) continue;
          // the holders of c around v are pairwise non-adjacent (they shared a colour): their moves do not interact
          for (auto &m : moves) { --load[c]; ++load[m.second]; color[m.first] = m.second; }
          --load[color[v]];
          color[v] = c;
          ++load[c];
          break;
        }
      }
      // close the gaps empty overflow colours leave
      std::vector<uint32_t> remap(load.size(), 0xffffffffu);
      uint32_t used = 0;
      for (uint32_t k = 0; k < load.size(); ++k)
        if (load[k] || k < K) remap[k] = used++;
      bool changed = false;
      for (uint32_t k = 0; k < load.size(); ++k) changed |= remap[k] != 0xffffffffu && remap[k] != k;
      if (changed) parallel_for(nvar, [&](uint64_t v0, uint64_t v1) { for (uint64_t v = v0; v < v1; ++v) color[v] = remap[color[v]]; });
      ncolor = 1;
      for (uint32_t k = 0; k < load.size(); ++k)
        if (load[k]) ncolor = std::max(ncolor, remap[k] + 1);
    }
  }
  return ncolor;
}

// Per-weight statistics of a learning sweep, from the CSR rows of a compiled image (the curvature-
// aware batch update of the CSR kernels and the epoch plan of the live-weight kernels read them).
static void class_statistics(const RawView &raw, GraphImage *g, uint64_t nvar_pad, uint32_t ncolor, uint64_t nweight,
                             uint64_t ow) {
  // ---- per colour class, per weight: how many SGD updates one learning launch makes (N) and
  // their summed curvature bound (Q), for the semi-implicit batch update (dwg_kernels.cuh).
  // q = (fv * sign range)^2 / 4 >= Var(f | rest): the conditional is a Bernoulli. ----
  {
    uint64_t cells = (uint64_t)ncolor * nweight;
    g->cls_shared = g->sell || cells > (64ull << 20);  // sell: whole-graph sums; too many cells: one class-averaged row
    uint64_t rows_n = g->cls_shared ? 1 : ncolor;
    g->cls_N_evid.assign(rows_n * nweight, 0.f); g->cls_Q_evid.assign(rows_n * nweight, 0.f);
    g->cls_N_all.assign(rows_n * nweight, 0.f); g->cls_Q_all.assign(rows_n * nweight, 0.f);
    g->cls_H_evid.assign(rows_n * nweight, 0.f); g->cls_H_all.assign(rows_n * nweight, 0.f);
    auto srange = [](uint32_t fc, uint64_t n) -> double {
      switch (fc) {
        case FC_AND: case FC_OR: case FC_EQUAL: case FC_IMPLY_NATURAL: return 2.0;
        case FC_LINEAR: return n > 1 ? (double)(n - 1) : 1.0;
        case FC_RATIO: return n > 1 ? std::log2((double)n) : 1.0;
        default: return 1.0;
      }
    };
    // delta(e) = |fv| * (sign range) of record e.  One SGD step of variable v moves weight k
    // by a gradient whose covariance over the draw is p(1-p) * delta delta^T <= 1/4 delta delta^T;
    // diag(|delta_k| * sum_j |delta_j|) dominates delta delta^T, so
    //   q(e) = 1/4 * delta(e) * sum_{e' learnable in the row} delta(e')
    // summed over a launch bounds the curvature of ALL weights moving together (the features
    // of one variable are perfectly correlated: a per-weight diagonal alone would overshoot).
    auto walk_row = [&](uint64_t r, auto &&fn) {
      uint64_t oc = g->has_multi ? g->row_oth[r] : 0;
      for (uint64_t e = g->row_ptr[r]; e < g->row_ptr[r + 1]; ++e) {
        uint8_t tag = g->e_tag[e];
        uint32_t form = tag >> TAG_FORM_SHIFT;
        uint64_t n = form != FORM_GENERAL ? form + 1 : g->oth[oc * ow];
        oc += form != FORM_GENERAL ? form : 1 + n;
        uint32_t w = g->e_wid[e];
        if (raw.w_fixed[g->w_perm[w]]) continue;
        double fv = g->fv_mode == FV_F32 ? g->e_fv32[e] : g->fv_mode == FV_F64 ? g->e_fv64[e] : 1.0;
        fn(w, std::fabs(fv) * srange(tag & TAG_FUNC_MASK, n));
      }
    };
    if (g->sell) {
      // Sliced-ELL graphs learn in ONE visit order over all classes (learn_visit below): whole-graph
      // sums, one row.  Accumulated as 2^-20 fixed point with integer atomics: the result does not
      // depend on the thread count or the interleaving.
      g->cls_shared = true;
      std::vector<uint64_t> acc(6 * nweight, 0);
      auto fx = [](double x) { return (uint64_t)std::llround(x * 1048576.0); };
      parallel_for(nvar_pad, [&](uint64_t d0, uint64_t d1) {
        for (uint64_t d = d0; d < d1; ++d) {
          uint64_t v = g->perm[d];
          if (v == 0xffffffffu) continue;
          const bool ev = g->var_is_evid[v];
          double rowsum = 0;
          walk_row(d, [&](uint32_t, double delta) { rowsum += delta; });
          walk_row(d, [&](uint32_t w, double delta) {
            const uint64_t q = fx(0.25 * delta * rowsum), hd = fx(0.25 * delta * delta);
            __atomic_fetch_add(&acc[0 * nweight + w], 1ull << 20, __ATOMIC_RELAXED);
            __atomic_fetch_add(&acc[1 * nweight + w], q, __ATOMIC_RELAXED);
            __atomic_fetch_add(&acc[2 * nweight + w], hd, __ATOMIC_RELAXED);
            if (ev) {
              __atomic_fetch_add(&acc[3 * nweight + w], 1ull << 20, __ATOMIC_RELAXED);
              __atomic_fetch_add(&acc[4 * nweight + w], q, __ATOMIC_RELAXED);
              __atomic_fetch_add(&acc[5 * nweight + w], hd, __ATOMIC_RELAXED);
            }
          });
        }
      });
      for (uint64_t w = 0; w < nweight; ++w) {
        g->cls_N_all[w] = (float)((double)acc[0 * nweight + w] / 1048576.0);
        g->cls_Q_all[w] = (float)((double)acc[1 * nweight + w] / 1048576.0);
        g->cls_H_all[w] = (float)((double)acc[2 * nweight + w] / 1048576.0);
        g->cls_N_evid[w] = (float)((double)acc[3 * nweight + w] / 1048576.0);
        g->cls_Q_evid[w] = (float)((double)acc[4 * nweight + w] / 1048576.0);
        g->cls_H_evid[w] = (float)((double)acc[5 * nweight + w] / 1048576.0);
      }
    } else {
    // every class owns its row of the tables: classes in parallel (the float sums of a row keep
    // their sequential order); a shared row is filled by one thread
    parallel_ranges(ncolor, g->cls_shared ? 1u : std::min<unsigned>(ncolor, host_threads()),
                    [&](uint64_t c0, uint64_t c1, unsigned) {
    for (uint64_t c = c0; c < c1; ++c) {
      uint64_t base = g->cls_shared ? 0 : (uint64_t)c * nweight;
      for (uint64_t d = g->color_start[c]; d < g->color_start[c + 1]; ++d) {
        uint64_t v = g->perm[d];
        if (v == 0xffffffffu) continue;
        bool cat = g->var_dtype[v], ev = g->var_is_evid[v];
        uint64_t r0 = g->has_categorical ? g->var_row[d] : d;
        uint64_t r = cat ? r0 + std::min<uint64_t>(g->evid_val[d], raw.var_card[v] ? raw.var_card[v] - 1 : 0) : r0;
        if (cat && raw.var_card[v] == 0) continue;
        double rowsum = 0;
        walk_row(r, [&](uint32_t, double delta) { rowsum += delta; });
        const float one = cat ? 2.f : 1.f;  // a categorical step also visits the row of the free value
        walk_row(r, [&](uint32_t w, double delta) {
          float q = (float)(0.25 * delta * rowsum) * one, hd = (float)(0.25 * delta * delta) * one;
          g->cls_N_all[base + w] += one; g->cls_Q_all[base + w] += q; g->cls_H_all[base + w] += hd;
          if (ev) { g->cls_N_evid[base + w] += one; g->cls_Q_evid[base + w] += q; g->cls_H_evid[base + w] += hd; }
        });
      }
    }
    }, 1);
    if (g->cls_shared && !g->sell)
      for (uint64_t w = 0; w < nweight; ++w) {
        g->cls_N_evid[w] /= ncolor; g->cls_Q_evid[w] /= ncolor; g->cls_H_evid[w] /= ncolor;
        g->cls_N_all[w] /= ncolor; g->cls_Q_all[w] /= ncolor; g->cls_H_all[w] /= ncolor;
      }
    }
  }

}

// ---- sliced-ELL blob of a Boolean graph (graph_image.h) from the CSR rows of a compiled image.
// wide[v]: variable v owns its slice(s) (warp per variable); wide_slices(v): how many. ----
template <typename WideSlices>
static std::string build_sell_blob(const RawView &raw, GraphImage *g, uint64_t nvar, uint64_t nvar_pad, uint32_t ncolor,
                                   const std::vector<uint8_t> &wide, WideSlices &&wide_slices, int *err) {
  // ---- sliced-ELL blob (graph_image.h): per 32-variable slice one contiguous run of words:
  // header, unary records (+ feature values), multi records (+ feature values), others.
  // A warp streams its slice with coalesced 128 B requests, every record exactly once. ----
  if (g->sell) {
    const uint64_t nslice = nvar_pad / 32;
    const uint64_t fvw = g->fv_mode == FV_F32 ? 1 : g->fv_mode == FV_F64 ? 2 : 0;  // words per value
    g->sl_off.assign(nslice + 1, 0);
    auto is_unary = [&](uint64_t e) { return (g->e_tag[e] >> TAG_FORM_SHIFT) == 0; };
    // slice shapes.  A slice whose first slot holds a wide variable is that variable's alone: record
    // k of a section goes to lane k % 32, column k / 32 (the lanes of a normal slice are variables)
    // slice -> first slice of the wide variable it belongs to (itself for a one-slice wide variable)
    std::vector<uint32_t> slice_hub(nslice, 0xffffffffu);
    parallel_for(nvar, [&](uint64_t v0, uint64_t v1) {
      for (uint64_t v = v0; v < v1; ++v)
        if (wide[v]) {
          const uint32_t s0 = g->inv[v] / 32, n = wide_slices(v);
          for (uint32_t q = 0; q < n; ++q) slice_hub[s0 + q] = s0;
        }
    });
    for (uint64_t sidx = 0; sidx < nslice; ++sidx)
      if (slice_hub[sidx] != 0xffffffffu && wide_slices(g->perm[(uint64_t)slice_hub[sidx] * 32]) > 1)
        g->part_slices.push_back((uint32_t)sidx);
    auto slice_is_wide = [&](uint64_t sidx) { return slice_hub[sidx] != 0xffffffffu; };
    for (uint64_t sidx = 0; sidx < nslice && !g->has_wide; ++sidx) g->has_wide = slice_is_wide(sidx);
    // the share of part q of n: records [q * per, min(total, (q + 1) * per)) of a section
    auto share = [](uint64_t total, uint32_t q, uint32_t n, uint64_t *lo, uint64_t *hi) {
      const uint64_t per = (total + n - 1) / n;
      *lo = std::min(total, per * q);
      *hi = std::min(total, per * (q + 1));
    };
    auto others_span = [&](uint64_t e, uint64_t o) -> uint64_t {
      const uint32_t form = g->e_tag[e] >> TAG_FORM_SHIFT;
      return form != FORM_GENERAL ? form : 1 + g->oth[o];
    };
    std::vector<uint32_t> s_un(nslice), s_mu(nslice), s_oc(nslice);
    uint64_t words = 0;
    parallel_for(nslice, [&](uint64_t a0, uint64_t a1) {
      for (uint64_t sidx = a0; sidx < a1; ++sidx) {
        uint32_t un = 0, mu = 0, oc = 0;
        if (slice_is_wide(sidx)) {
          const uint64_t d = (uint64_t)slice_hub[sidx] * 32;
          const uint32_t q = (uint32_t)(sidx - slice_hub[sidx]), n = wide_slices(g->perm[d]);
          uint64_t U = 0;
          for (uint64_t e = g->row_ptr[d]; e < g->row_ptr[d + 1]; ++e) U += is_unary(e);
          const uint64_t M = g->row_ptr[d + 1] - g->row_ptr[d] - U;
          uint64_t ulo, uhi, mlo, mhi;
          share(U, q, n, &ulo, &uhi);
          share(M, q, n, &mlo, &mhi);
          uint32_t lane_oc[32] = {0};
          uint64_t m = 0, o = g->has_multi ? g->row_oth[d] : 0;
          for (uint64_t e = g->row_ptr[d]; e < g->row_ptr[d + 1]; ++e) {
            if (is_unary(e)) continue;
            const uint64_t span = others_span(e, o);
            if (m >= mlo && m < mhi) lane_oc[(m - mlo) % 32] += (uint32_t)span;
            o += span;
            ++m;
          }
          un = (uint32_t)((uhi - ulo + 31) / 32); mu = (uint32_t)((mhi - mlo + 31) / 32);
          for (uint32_t l = 0; l < 32; ++l) oc = std::max(oc, lane_oc[l]);
        } else {
          for (uint64_t d = sidx * 32; d < sidx * 32 + 32; ++d) {
            uint32_t u = 0;
            for (uint64_t e = g->row_ptr[d]; e < g->row_ptr[d + 1]; ++e) u += is_unary(e);
            un = std::max(un, u);
            mu = std::max<uint32_t>(mu, g->row_ptr[d + 1] - g->row_ptr[d] - u);
            if (g->has_multi) oc = std::max(oc, g->row_oth[d + 1] - g->row_oth[d]);
          }
        }
        s_un[sidx] = un; s_mu[sidx] = mu; s_oc[sidx] = oc;
      }
    });
    for (uint64_t sidx = 0; sidx < nslice; ++sidx) {
      const uint32_t un = s_un[sidx], mu = s_mu[sidx], oc = s_oc[sidx];
      g->sl_off[sidx] = (uint32_t)(words / 8);
      uint64_t w = SH_WORDS + 32ull * ((uint64_t)(un + mu) * (1 + fvw) + oc);
      g->sell_max_words = (uint32_t)std::max<uint64_t>(g->sell_max_words, std::min<uint64_t>(w, U32_LIMIT));
      words += w;
      if (words / 8 >= U32_LIMIT) { *err = DWG_ERR_UNSUPPORTED; return "sliced-ELL stream exceeds 2^32 x 32 B"; }
    }
    g->sl_off[nslice] = (uint32_t)(words / 8);
    g->blob.assign(words, 0);
    {  // learning order: by id block of the slice's first variable, then by class, then by position
      std::vector<uint64_t> vk(nslice);
      parallel_for(ncolor, [&](uint64_t c0, uint64_t c1) {
        for (uint64_t c = c0; c < c1; ++c)
          for (uint64_t sidx = g->color_start[c] / 32; sidx < g->color_start[c + 1] / 32; ++sidx) {
            uint64_t v = 0;
            if (slice_hub[sidx] != 0xffffffffu) v = g->perm[(uint64_t)slice_hub[sidx] * 32];  // every part of a hub: the hub
            else
              for (uint64_t d = sidx * 32; d < sidx * 32 + 32; ++d)
                if (g->perm[d] != 0xffffffffu) { v = g->perm[d]; break; }
            vk[sidx] = ((v / g->id_block) << 40) | (c << 24 & 0xffff000000ull) | (sidx & 0xffffffu);
          }
      });
      // (the low fields only break ties; the slice index is recovered from a parallel array)
      g->learn_visit.resize(nslice);
      std::iota(g->learn_visit.begin(), g->learn_visit.end(), 0u);
      std::stable_sort(g->learn_visit.begin(), g->learn_visit.end(), [&](uint32_t a, uint32_t b) {
        const uint64_t ka = vk[a] >> 24, kb = vk[b] >> 24;
        return ka != kb ? ka < kb : a < b;
      });
    }
    // sign tables (host mirror of dev_sign, dwg_device.cuh): codes 0 -> -1, 1 -> 0, 2 -> +1
    auto sign2 = [](uint32_t fc, uint32_t nsb, bool head) -> int {  // arity 2
      const bool body_all = nsb == 1;
      switch (fc) {
        case FC_AND: return (body_all && head) ? 1 : -1;
        case FC_AND_CATEGORICAL: return (body_all && head) ? 1 : 0;
        case FC_OR: return (nsb > 0 || head) ? 1 : -1;
        case FC_EQUAL: return (nsb + (head ? 1u : 0u)) != 1 ? 1 : -1;
        case FC_IMPLY_NATURAL: return body_all ? (head ? 1 : -1) : 0;
        case FC_IMPLY_MLN: return body_all ? (head ? 1 : 0) : 1;
        default: return (head || nsb == 0) ? 1 : 0;  // LINEAR / RATIO (log2(1+1) = 1) / LOGICAL
      }
    };
    auto code_mismatch = [](uint32_t fc) -> uint32_t {  // arity 1, variable does not satisfy
      switch (fc) {
        case FC_AND: case FC_OR: case FC_IMPLY_NATURAL: return 0;  // -1
        case FC_EQUAL: return 2;                                   // +1
        default: return 1;                                         //  0
      }
    };
    auto put_fv = [&](uint32_t *slot, uint64_t e) {
      if (g->fv_mode == FV_F32) std::memcpy(slot, &g->e_fv32[e], 4);
      else if (g->fv_mode == FV_F64) std::memcpy(slot, &g->e_fv64[e], 8);
    };
    parallel_for(nslice, [&](uint64_t a0, uint64_t a1) {
    for (uint64_t sidx = a0; sidx < a1; ++sidx) {
      uint32_t *h = g->blob.data() + (uint64_t)g->sl_off[sidx] * 8;
      const uint32_t un = s_un[sidx], mu = s_mu[sidx], oc = s_oc[sidx];
      const uint64_t o_un = SH_WORDS, o_unfv = o_un + 32ull * un, o_mu = o_unfv + 32ull * un * fvw,
                     o_mufv = o_mu + 32ull * mu, o_oth = o_mufv + 32ull * mu * fvw;
      const bool is_wide = slice_is_wide(sidx);
      const uint32_t nvars = is_wide ? 1 : 32;  // variables whose rows the slice holds
      const uint64_t dbase = is_wide ? (uint64_t)slice_hub[sidx] * 32 : sidx * 32;  // slot of lane 0's variable
      const uint32_t pq = is_wide ? (uint32_t)(sidx - slice_hub[sidx]) : 0, pn = is_wide ? wide_slices(g->perm[dbase]) : 1;
      uint64_t ulo = 0, uhi = ~0ull, mlo = 0, mhi = ~0ull;  // this part's share of the variable's records
      if (pn > 1) {
        uint64_t U = 0;
        for (uint64_t e = g->row_ptr[dbase]; e < g->row_ptr[dbase + 1]; ++e) U += is_unary(e);
        share(U, pq, pn, &ulo, &uhi);
        share(g->row_ptr[dbase + 1] - g->row_ptr[dbase] - U, pq, pn, &mlo, &mhi);
      }
      // is every multi record of the slice a one-other-variable record?  (a hub: of the whole hub,
      // so that all its parts are read the same way)
      bool binary = mu > 0 || pn > 1, antisym = true;
      for (uint64_t d = dbase; d < dbase + nvars && binary; ++d) {
        for (uint64_t e = g->row_ptr[d]; e < g->row_ptr[d + 1]; ++e)
          if (!is_unary(e) && (g->e_tag[e] >> TAG_FORM_SHIFT) != 1) binary = false;
      }
      if (pn > 1 && mu == 0) binary = false;
      h[SH_N_UNARY] = un; h[SH_N_MULTI] = mu; h[SH_N_OTH] = oc;
      h[SH_SELF_OFF] = g->sl_off[sidx];
      uint32_t ju[32] = {0}, jm[32] = {0}, jo[32] = {0};  // per lane: unary / multi columns and others words written
      for (uint32_t l = 0; l < 32; ++l) {
        // (every part of a hub carries the hub's bits at lane 0; the other lanes are padding slots)
        const uint64_t d = l == 0 ? dbase : sidx * 32 + l;
        if (g->var_flags[d] & VF_PAD) h[SH_PAD_MASK] |= 1u << l;
        if (g->var_flags[d] & VF_EVID) h[SH_EVID_MASK] |= 1u << l;
        if (g->evid_val[d] == 1) h[SH_EVVAL_MASK] |= 1u << l;
        if (l >= nvars) continue;
        uint64_t ku = 0, km = 0;  // records of this variable seen so far
        uint64_t o = g->has_multi ? g->row_oth[d] : 0;
        for (uint64_t e = g->row_ptr[d]; e < g->row_ptr[d + 1]; ++e) {
          const uint8_t tag = g->e_tag[e];
          const uint32_t fc = tag & TAG_FUNC_MASK, form = tag >> TAG_FORM_SHIFT;
          const uint32_t wid = g->e_wid[e];
          const uint32_t base = wid | (raw.w_fixed[g->w_perm[wid]] ? SELL_FIXED_BIT : 0u);
          if (form == 0) {
            if (ku < ulo || ku >= uhi) { ++ku; continue; }  // another part's record
            const uint32_t tl = is_wide ? (uint32_t)((ku - ulo) % 32) : l;  // target lane
            const uint32_t mm = code_mismatch(fc);
            const uint32_t s1 = (tag & TAG_SELF_EQ1) ? 2 : mm, s0 = (tag & TAG_SELF_EQ1) ? mm : 2;
            if (s1 + s0 != 2) antisym = false;
            h[o_un + 32ull * ju[tl] + tl] = sell_unary_record(wid, raw.w_fixed[g->w_perm[wid]], s1, s0);
            put_fv(h + o_unfv + (32ull * ju[tl] + tl) * fvw, e);
            ++ju[tl]; ++ku;
            continue;
          }
          const uint64_t span = others_span(e, o);
          if (km < mlo || km >= mhi) { ++km; o += span; continue; }  // another part's record
          const uint32_t tl = is_wide ? (uint32_t)((km - mlo) % 32) : l;
          uint32_t t = tag;
          if (binary) {
            // sign table over (this variable's value x, other variable satisfied)
            const uint32_t self_eq = (tag & TAG_SELF_EQ1) ? 1 : 0;
            const bool self_head = tag & TAG_SELF_HEAD;
            t = 0;
            for (uint32_t x = 0; x < 2; ++x)
              for (uint32_t sat = 0; sat < 2; ++sat) {
                const bool m = x == self_eq;
                const int sgn = self_head ? sign2(fc, sat, m) : sign2(fc, m ? 1u : 0u, sat != 0);
                t |= (uint32_t)(sgn + 1) << (2 * (2 * x + sat));
              }
            h[o_oth + 32ull * jm[tl] + tl] = g->oth[o];  // record j owns others column j
          } else {
            for (uint64_t k = 0; k < span; ++k) h[o_oth + 32ull * (jo[tl] + k) + tl] = g->oth[o + k];
            jo[tl] += (uint32_t)span;
          }
          o += span;
          h[o_mu + 32ull * jm[tl] + tl] = base | (t << 24);
          put_fv(h + o_mufv + (32ull * jm[tl] + tl) * fvw, e);
          ++jm[tl]; ++km;
        }
      }
      for (uint32_t l = 0; l < 32; ++l) {
        for (uint32_t j = ju[l]; j < un; ++j) h[o_un + 32ull * j + l] = SELL_UNARY_NOP;
        for (uint32_t j = jm[l]; j < mu; ++j) h[o_mu + 32ull * j + l] = binary ? SELL_BINARY_NOP << 24 : 0u;  // form 0
      }
      h[SH_FLAGS] = (antisym ? SF_ANTISYM : 0u) | (binary ? SF_BINARY : 0u) | (is_wide ? SF_WIDE : 0u) |
                    (pn > 1 ? SF_PART | (pq << 4) | (pn << 18) : 0u);
    }
    });
  }
  return "";
}

// ---- categorical sliced-ELL blob (graph_image.h CELL_*) from the CSR rows of a compiled image:
// slice shapes, offsets, learning order, and per slice header | meta | first tally row | record
// columns (rows in value order, a row's records in factor-id order) | others columns.  Clears
// g->cell when some variable has more two-variable factors than a lane's mask holds. ----
static std::string build_cell_blob(GraphImage *g, uint64_t nvar_pad, uint32_t ncolor, int *err) {
  std::vector<uint32_t> cell_rec, cell_oc;
  // host mirror of dev_sign (dwg_device.cuh) for one- and two-variable factors, as sign codes
    auto code1 = [](uint32_t fc, bool self_sat) -> uint32_t {  // arity 1
      if (self_sat) return 2;
      switch (fc) {
        case FC_AND: case FC_OR: case FC_IMPLY_NATURAL: return 0;
        case FC_EQUAL: return 2;
        default: return 1;
      }
    };
    auto code2 = [](uint32_t fc, uint32_t nsb, bool head) -> uint32_t {  // arity 2: one body position
      const bool body_all = nsb == 1;
      int sgn;
      switch (fc) {
        case FC_AND: sgn = (body_all && head) ? 1 : -1; break;
        case FC_AND_CATEGORICAL: sgn = (body_all && head) ? 1 : 0; break;
        case FC_OR: sgn = (nsb > 0 || head) ? 1 : -1; break;
        case FC_EQUAL: sgn = (nsb + (head ? 1u : 0u)) != 1 ? 1 : -1; break;
        case FC_IMPLY_NATURAL: sgn = body_all ? (head ? 1 : -1) : 0; break;
        case FC_IMPLY_MLN: sgn = body_all ? (head ? 1 : 0) : 1; break;
        default: sgn = (head || nsb == 0) ? 1 : 0; break;  // LINEAR / RATIO (log2(1+1) = 1) / LOGICAL
      }
      return (uint32_t)(sgn + 1);
    };
  if (g->cell) {
    const uint64_t nslice = nvar_pad / 32;
    // slice shapes: record columns (an empty row still holds one record), others columns
    cell_rec.assign(nslice, 0); cell_oc.assign(nslice, 0);
    std::vector<uint32_t> &s_rec = cell_rec, &s_oc = cell_oc;
    int too_many_others = 0;
    parallel_for(nslice, [&](uint64_t a0, uint64_t a1) {
      for (uint64_t sidx = a0; sidx < a1; ++sidx) {
        uint32_t rec = 0, oc = 0;
        for (uint64_t d = sidx * 32; d < sidx * 32 + 32; ++d) {
          if (g->var_flags[d] & VF_PAD) continue;
          uint32_t r = 0, o = 0;
          for (uint64_t row = g->var_row[d]; row < g->var_row[d + 1]; ++row) {
            const uint32_t n = g->row_ptr[row + 1] - g->row_ptr[row];
            r += n ? n : 1u;
            if (g->has_multi) o += g->row_oth[row + 1] - g->row_oth[row];
          }
          rec = std::max(rec, r); oc = std::max(oc, o);
        }
        s_rec[sidx] = rec; s_oc[sidx] = oc;
        if (oc > CELL_MAX_OTH) __atomic_store_n(&too_many_others, 1, __ATOMIC_RELAXED);
      }
    });
    if (too_many_others) g->cell = false;  // a variable with more than 32 two-variable factors: the CSR kernels sweep this graph
  }
  if (g->cell) {
    const uint64_t nslice = nvar_pad / 32;
    std::vector<uint32_t> &s_rec = cell_rec, &s_oc = cell_oc;
    g->sl_off.assign(nslice + 1, 0);
    uint64_t words = 0;
    for (uint64_t sidx = 0; sidx < nslice; ++sidx) {
      g->sl_off[sidx] = (uint32_t)(words / 8);
      const uint64_t w = SH_WORDS + 32ull * (2 + s_rec[sidx] + s_oc[sidx]);
      g->sell_max_words = (uint32_t)std::max<uint64_t>(g->sell_max_words, w);
      words += w;
      if (words / 8 >= U32_LIMIT) { *err = DWG_ERR_UNSUPPORTED; return "categorical sliced-ELL stream exceeds 2^32 x 32 B"; }
    }
    g->sl_off[nslice] = (uint32_t)(words / 8);
    g->blob.assign(words, 0);
    {  // learning order (k_cell_learn_live): by id block of the slice's first variable, then by class, then by position
      std::vector<uint64_t> vk(nslice);
      parallel_for(ncolor, [&](uint64_t c0, uint64_t c1) {
        for (uint64_t c = c0; c < c1; ++c)
          for (uint64_t sidx = g->color_start[c] / 32; sidx < g->color_start[c + 1] / 32; ++sidx) {
            uint64_t v = 0;
            for (uint64_t d = sidx * 32; d < sidx * 32 + 32; ++d)
              if (g->perm[d] != 0xffffffffu) { v = g->perm[d]; break; }
            vk[sidx] = ((v / g->id_block) << 16) | (c & 0xffffu);
          }
      });
      g->learn_visit.resize(nslice);
      std::iota(g->learn_visit.begin(), g->learn_visit.end(), 0u);
      std::stable_sort(g->learn_visit.begin(), g->learn_visit.end(), [&](uint32_t a, uint32_t b) {
        return vk[a] != vk[b] ? vk[a] < vk[b] : a < b;
      });
    }
    parallel_for(nslice, [&](uint64_t a0, uint64_t a1) {
      for (uint64_t sidx = a0; sidx < a1; ++sidx) {
        uint32_t *h = g->blob.data() + (uint64_t)g->sl_off[sidx] * 8;
        const uint32_t nrec = s_rec[sidx], noc = s_oc[sidx];
        const uint64_t o_meta = SH_WORDS, o_row0 = o_meta + 32, o_rec = o_row0 + 32, o_oth = o_rec + 32ull * nrec;
        h[CH_N_REC] = nrec; h[CH_N_OTH] = noc; h[CH_SELF_OFF] = g->sl_off[sidx];
        uint32_t kmax = 0;
        for (uint32_t l = 0; l < 32; ++l) {
          const uint64_t d = sidx * 32 + l;
          uint32_t jr = 0, jo = 0;
          if (g->var_flags[d] & VF_PAD) {
            h[CH_PAD_MASK] |= 1u << l;
            h[o_row0 + l] = g->var_row[d];
          } else {
            if (g->var_flags[d] & VF_EVID) h[CH_EVID_MASK] |= 1u << l;
            const uint32_t K = g->var_row[d + 1] - g->var_row[d];
            kmax = std::max(kmax, K);
            h[o_meta + l] = K | (g->evid_val[d] << 8);
            h[o_row0 + l] = g->var_row[d];
            for (uint64_t row = g->var_row[d]; row < g->var_row[d + 1]; ++row) {
              bool first = true;
              uint64_t o = g->has_multi ? g->row_oth[row] : 0;
              for (uint64_t e = g->row_ptr[row]; e < g->row_ptr[row + 1]; ++e) {
                const uint8_t tag = g->e_tag[e];
                const uint32_t fc = tag & TAG_FUNC_MASK, form = tag >> TAG_FORM_SHIFT;
                uint32_t table = 0;
                for (uint32_t ss = 0; ss < 2; ++ss)
                  for (uint32_t os = 0; os < 2; ++os) {
                    uint32_t c;
                    if (form == 0) c = code1(fc, ss);
                    else if (tag & TAG_SELF_HEAD) c = code2(fc, os, ss);
                    else c = code2(fc, ss, os);
                    table |= c << (2 * (2 * ss + os));
                  }
                uint32_t rec = g->e_wid[e] | (table << 22) | (first ? CELL_ROW_BIT : 0u);
                if (form == 1) {
                  rec |= CELL_MULTI_BIT;
                  h[o_oth + 32ull * jo + l] = (g->oth[2 * o] << 8) | (g->oth[2 * o + 1] & 0xffu);
                  ++jo; ++o;
                }
                h[o_rec + 32ull * jr + l] = rec;
                ++jr; first = false;
              }
              if (first) { h[o_rec + 32ull * jr + l] = CELL_NOP_ROW; ++jr; }  // a value without factors
            }
          }
          for (; jr < nrec; ++jr) h[o_rec + 32ull * jr + l] = CELL_NOP;
        }
        h[CH_KMAX] = kmax;
      }
    });
  }
  return "";
}

std::string compile_graph(const RawView &raw, int schedule, GraphImage *g, int *err) {
  *err = DWG_ERR_INVALID;
  *g = GraphImage();
  PhaseTimer pt;
  const uint64_t nvar = raw.nvar, nfac = raw.nfac, nweight = raw.nweight;
  const uint64_t nedge = nfac ? raw.fac_ptr[nfac] : 0;
  if (nvar >= U32_LIMIT / 2 || nfac >= U32_LIMIT || nweight >= U32_LIMIT) {
    *err = DWG_ERR_UNSUPPORTED;
    return "graph exceeds the 32-bit device layout (variables < 2^31, factors/weights < 2^32)";
  }
  g->nvar = nvar; g->nfac = nfac; g->nweight = nweight; g->nedge = nedge;
  g->w_init.assign(raw.w_init, raw.w_init + nweight);
  g->w_fixed.assign(raw.w_fixed, raw.w_fixed + nweight);

  // ---- variables (src/binary_format.cc:72-126) ----
  std::vector<uint64_t> adense(nvar);
  g->var_dtype.resize(nvar); g->var_is_evid.resize(nvar);
  uint64_t max_card = 2;
  for (uint64_t v = 0; v < nvar; ++v) {
    if (raw.var_dtype[v] > 1)
      return "variable " + std::to_string(v) + ": only Boolean and categorical variables are supported";
    bool cat = raw.var_dtype[v] == 1;
    g->var_dtype[v] = cat;
    bool ev = raw.var_role[v] >= 1;  // :107
    g->var_is_evid[v] = ev;
    adense[v] = ev ? raw.var_init[v] : 0;  // :109
    if (ev) ++g->nvar_evid; else ++g->nvar_query;
    if (cat) {
      g->has_categorical = true;
      if (raw.var_card[v] >= U32_LIMIT) { *err = DWG_ERR_UNSUPPORTED; return "cardinality >= 2^32"; }
      max_card = std::max<uint64_t>(max_card, raw.var_card[v]);
    }
  }

  // ---- domains (src/binary_format.cc:192-226) ----
  Domains dom;
  dom.block_of.assign(nvar, -1);
  dom.tab_ptr.assign(raw.ndom + 1, 0);
  for (uint64_t b = 0; b < raw.ndom; ++b) {
    uint64_t v = raw.dom_vid[b], n = raw.dom_ptr[b + 1] - raw.dom_ptr[b];
    if (v >= nvar) return "domain block for unknown variable " + std::to_string(v);
    if (!g->var_dtype[v]) return "domain given for Boolean variable " + std::to_string(v);  // :206
    if (raw.var_card[v] != n) return "domain size != cardinality for variable " + std::to_string(v);  // :207
    dom.block_of[v] = (int64_t)b;
    dom.tab_ptr[b + 1] = dom.tab_ptr[b] + n;
  }
  dom.tab.resize(dom.tab_ptr[raw.ndom]);
  for (uint64_t b = 0; b < raw.ndom; ++b) {
    DomEntry *t = dom.tab.data() + dom.tab_ptr[b];
    uint64_t n = raw.dom_ptr[b + 1] - raw.dom_ptr[b];
    for (uint64_t i = 0; i < n; ++i) t[i] = {raw.dom_val[raw.dom_ptr[b] + i], (uint32_t)i};
    std::stable_sort(t, t + n, [](const DomEntry &a, const DomEntry &c) { return a.val < c.val; });
    // last occurrence of a repeated value wins: keep, for equal values, the largest index first
    for (uint64_t i = 0; i < n;) {
      uint64_t j = i;
      while (j + 1 < n && t[j + 1].val == t[i].val) ++j;
      if (j > i) std::swap(t[i], t[j]);
      i = j + 1;
    }
  }
  for (uint64_t v = 0; v < nvar; ++v) {
    // evidence value remapped only when non-zero (:220-223) -- sparse value 0 stays dense 0
    if (dom.block_of[v] >= 0 && adense[v]) {
      adense[v] = dom.lookup(v, adense[v]);
      if (adense[v] == INVALID)
        return "initial value of variable " + std::to_string(v) + " is not in its domain";
    }
  }

  pt.lap("variables + domains");
  // ---- factors (src/binary_format.cc:128-190) ----
  BulkVec<uint32_t> edense(nedge);  // FactorToVariable::dense_equal_to (every entry written below)
  std::vector<uint32_t> deg(nvar + 1, 0);
  bool coupled = false, eq_is_bit = true;
  {
    // factor ranges in parallel; the error reported is the one of the lowest factor id, as a
    // sequential scan would find it
    struct Part {
      uint64_t bad_f = INVALID;
      std::string msg;
      int code = DWG_ERR_INVALID;
      bool coupled = false, eq_is_bit = true, multi = false;
    };
    const unsigned T = host_threads();
    std::vector<Part> parts(T);
    parallel_ranges(nfac, T, [&](uint64_t f0, uint64_t f1, unsigned p) {
      Part &P = parts[p];
      auto bad = [&](uint64_t f, std::string m, int code = DWG_ERR_INVALID) {
        P.bad_f = f; P.msg = std::move(m); P.code = code;
      };
      for (uint64_t f = f0; f < f1; ++f) {
        if (func_code(raw.fac_func[f]) < 0)
          return bad(f, "factor " + std::to_string(f) + ": unsupported FACTOR_FUNCTION_TYPE = " +
                            std::to_string(raw.fac_func[f]));
        uint64_t a = raw.fac_ptr[f], b = raw.fac_ptr[f + 1];
        // (a zero-arity factor is legal input: it sits in no adjacency list and is never evaluated)
        if (raw.fac_wid[f] >= nweight) return bad(f, "factor " + std::to_string(f) + " refers to an unknown weight");
        if (b - a > 1) P.multi = true;
        for (uint64_t e = a; e < b; ++e) {
          uint64_t v = raw.edge_vid[e];
          if (v >= nvar) return bad(f, "factor " + std::to_string(f) + " refers to an unknown variable");  // :159
          uint64_t d = dom.lookup(v, raw.edge_eq[e]);
          if (d == INVALID)
            return bad(f, "factor " + std::to_string(f) + ": value " + std::to_string(raw.edge_eq[e]) +
                              " is not in the domain of variable " + std::to_string(v));
          if (d >= U32_LIMIT) return bad(f, "equal_to value >= 2^32", DWG_ERR_UNSUPPORTED);
          edense[e] = (uint32_t)d;
          if (d > 1) P.eq_is_bit = false;
          if (v != raw.edge_vid[a]) P.coupled = true;
          // the reference asserts value_dense < cardinality in construct_index (:186)
          if (g->var_dtype[v] && d >= raw.var_card[v])
            return bad(f, "factor " + std::to_string(f) + ": value index out of range for variable " +
                              std::to_string(v));
          __atomic_fetch_add(&deg[v + 1], 1u, __ATOMIC_RELAXED);
        }
      }
    });
    const Part *first_bad = nullptr;
    for (const Part &P : parts) {
      if (P.bad_f != INVALID && (!first_bad || P.bad_f < first_bad->bad_f)) first_bad = &P;
      coupled |= P.coupled; eq_is_bit &= P.eq_is_bit;
      if (P.multi) g->has_multi = true;
    }
    if (first_bad) { *err = first_bad->code; return first_bad->msg; }
  }
  for (uint64_t v = 0; v < nvar; ++v) deg[v + 1] += deg[v];

  pt.lap("factors");
  // ---- adjacency + construct_index (src/factor_graph.cc:131-199) ----
  BulkVec<Inc> inc(nedge);
  {
    std::vector<uint32_t> fill(deg.begin(), deg.end() - 1);
    parallel_for(nfac, [&](uint64_t f0, uint64_t f1) {
      for (uint64_t f = f0; f < f1; ++f)
        for (uint64_t e = raw.fac_ptr[f]; e < raw.fac_ptr[f + 1]; ++e) {
          uint64_t v = raw.edge_vid[e];
          inc[__atomic_fetch_add(&fill[v], 1u, __ATOMIC_RELAXED)] = {g->var_dtype[v] ? edense[e] : 0u, (uint32_t)f};
        }
    });
  }
  // per variable: entries sorted by (value, factor id) -- the parallel fill leaves a list in
  // arrival order, and equal (value, factor) entries are interchangeable -- then consecutive
  // duplicates dropped (same factor under the same value, :179-195)
  std::vector<uint32_t> inc_end(nvar);
  parallel_for(nvar, [&](uint64_t v0, uint64_t v1) {
    for (uint64_t v = v0; v < v1; ++v) {
      Inc *a = inc.data() + deg[v], *b = inc.data() + deg[v + 1];
      std::sort(a, b, [](const Inc &x, const Inc &y) { return x.key != y.key ? x.key < y.key : x.fid < y.fid; });
      Inc *o = a;
      for (Inc *q = a; q != b; ++q)
        if (o == a || o[-1].key != q->key || o[-1].fid != q->fid) *o++ = *q;
      inc_end[v] = (uint32_t)(o - inc.data());
    }
  });

  pt.lap("adjacency + index");
  // ---- rows in ORIGINAL order (var_val_base, values, truthiness) ----
  g->value_base.resize(nvar + 1);
  uint64_t nrow = 0;
  for (uint64_t v = 0; v < nvar; ++v) {
    g->value_base[v] = nrow;
    nrow += g->var_dtype[v] ? raw.var_card[v] : 1;
  }
  g->value_base[nvar] = nrow;
  if (nrow >= U32_LIMIT) { *err = DWG_ERR_UNSUPPORTED; return "more than 2^32 (variable, value) rows"; }
  g->nrow = nrow;
  g->values.assign(nrow, 0);
  std::vector<double> truth_orig(nrow, 0.0);
  std::vector<uint8_t> has_truth(nvar, 0);
  for (uint64_t v = 0; v < nvar; ++v) {
    if (!g->var_dtype[v]) continue;
    uint64_t base = g->value_base[v], K = raw.var_card[v];
    int64_t b = dom.block_of[v];
    if (b < 0) {
      for (uint64_t j = 0; j < K; ++j) g->values[base + j] = j;  // :169-172
      continue;
    }
    // recover the list from the map (:151-164): slots of overwritten duplicates stay INVALID
    for (uint64_t j = 0; j < K; ++j) g->values[base + j] = INVALID;
    double total = 0;
    const DomEntry *t = dom.tab.data() + dom.tab_ptr[b];
    for (uint64_t i = 0; i < K; ++i) {
      if (i > 0 && t[i].val == t[i - 1].val) continue;  // not the winning occurrence
      uint64_t j = t[i].idx;
      g->values[base + j] = t[i].val;
      truth_orig[base + j] = raw.dom_truth[raw.dom_ptr[b] + j];
      total += truth_orig[base + j];
    }
    // has_truthiness(): !is_linear_zero(total), src/variable.h:105, src/common.h:111
    has_truth[v] = !(total <= 0.000001 && total >= -0.000001);
    if (has_truth[v]) g->has_truthiness = true;
  }

  pt.lap("rows");
  // ---- schedule ----
  const bool auto_schedule = schedule == DWG_SCHED_AUTO;
  if (schedule == DWG_SCHED_AUTO) {
    uint64_t nrec = 0;
    for (uint64_t v = 0; v < nvar; ++v) nrec += inc_end[v] - deg[v];
    schedule = nrec <= DWG_SEQUENTIAL_MAX_EDGES ? DWG_SCHED_SEQUENTIAL
               : coupled ? DWG_SCHED_CHROMATIC : DWG_SCHED_HOGWILD;
  }
  if (schedule != DWG_SCHED_HOGWILD && schedule != DWG_SCHED_CHROMATIC && schedule != DWG_SCHED_SEQUENTIAL)
    return "bad schedule";
  std::vector<uint32_t> color(nvar, 0);
  uint32_t ncolor = 1;
  if (schedule == DWG_SCHED_CHROMATIC && coupled) {
    // greedy colouring in id order of the conflict graph "share a factor"
    double cost = 0;
    for (uint64_t f = 0; f < nfac; ++f) {
      double n = (double)(raw.fac_ptr[f + 1] - raw.fac_ptr[f]);
      cost += n * n;
    }
    // the conflict lists below hold sum n (n - 1) entries: beyond ~4 G of them (16 GB) colouring is
    // not worth its memory.  AUTO then sweeps Hogwild like the reference's threads; an explicit
    // chromatic request is refused.
    if (cost > 4e9) {
      if (!auto_schedule) {
        *err = DWG_ERR_UNSUPPORTED;
        return "chromatic schedule: factor arities too large to colour; use the hogwild schedule";
      }
      schedule = DWG_SCHED_HOGWILD;
    }
  }
  if (schedule == DWG_SCHED_CHROMATIC && coupled) {
    ncolor = colour_variables(raw, nvar, inc, deg, inc_end, color);
  } else if (schedule == DWG_SCHED_CHROMATIC) {
    // nothing couples two variables: one class, i.e. the hogwild launch, and it is exact
  }
  g->schedule = schedule;

  pt.lap("schedule (colouring)");
  // ---- layout knobs ----
  g->assign_bytes = max_card <= 256 ? 1 : 4;
  g->oth_packed = !g->has_categorical && eq_is_bit;
  g->fv_mode = FV_ONE;
  for (uint64_t f = 0; f < nfac; ++f) {
    double x = raw.fac_fv[f];
    if (x == 1.0) continue;
    if ((double)(float)x == x) g->fv_mode = std::max(g->fv_mode, (int)FV_F32);
    else g->fv_mode = FV_F64;
  }
  // Boolean-only graphs under a parallel schedule get the sliced-ELL stream (graph_image.h)
  // (a Boolean edge with equal_to >= 2 -- legal input, never satisfied -- needs two-word others
  // entries, which only the CSR kernels read)
  g->sell = !g->has_categorical && g->oth_packed && schedule != DWG_SCHED_SEQUENTIAL && nweight <= SELL_MAX_WEIGHTS;
  for (uint64_t v = 0; v < nvar && g->sell; ++v)
    if (adense[v] > 1) g->sell = false;  // a Boolean evidence value outside {0,1}: CSR kernels only

  // how one edge record of variable v for factor f is encoded (shared by the sort key, the CSR
  // emission and the byte accounting): compact forms need v exactly once among <= 3 positions
  auto edge_form = [&](uint64_t v, uint32_t f, uint64_t k, bool cat, uint64_t *self_pos_out) -> int {
    uint64_t a = raw.fac_ptr[f], b = raw.fac_ptr[f + 1], n = b - a;
    uint64_t nself = 0, self_pos = 0;
    for (uint64_t e = a; e < b; ++e)
      if (raw.edge_vid[e] == v) { ++nself; self_pos = e - a; }
    uint32_t self_eq = edense[a + self_pos];
    if (self_pos_out) *self_pos_out = self_pos;
    bool compact = nself == 1 && n <= 3 && (cat ? self_eq == k : self_eq <= 1);
    return compact ? (int)(n - 1) : (int)FORM_GENERAL;
  };
  // ... evaluated ONCE per incidence (a random read of the factor each time), kept as
  // form | self position << 2 for the three phases that need it (sort keys, sizing, CSR fill)
  BulkVec<uint8_t> finfo(nedge);
  int general_seen = 0, wide_seen = 0;
  parallel_for(nvar, [&](uint64_t v0, uint64_t v1) {
    bool any_general = false, any_wide = false;
    for (uint64_t v = v0; v < v1; ++v) {
      const bool cat = g->var_dtype[v];
      for (uint32_t i = deg[v]; i < inc_end[v]; ++i) {
        uint64_t self_pos = 0;
        const int form = edge_form(v, inc[i].fid, inc[i].key, cat, &self_pos);
        finfo[i] = (uint8_t)(form | ((form != (int)FORM_GENERAL ? self_pos : 0) << 2));
        any_general |= form == (int)FORM_GENERAL;
        any_wide |= form >= 2;
      }
      if (any_general) __atomic_store_n(&general_seen, 1, __ATOMIC_RELAXED);
      if (any_wide) __atomic_store_n(&wide_seen, 1, __ATOMIC_RELAXED);
    }
  });
  // a general-form record walks the others stream even when its factor has one variable (a Boolean
  // edge with equal_to >= 2, a variable twice in a factor): the stream and its row offsets must exist
  if (general_seen) g->has_multi = true;
  // graphs of categorical variables only, factors over at most two variables, feature values 1: the
  // categorical sliced-ELL stream (graph_image.h CELL_*); everything else keeps the CSR kernels
  {
    bool all_cat = nvar > 0;
    for (uint64_t v = 0; v < nvar && all_cat; ++v) all_cat = g->var_dtype[v] && raw.var_card[v] >= 1;
    g->cell = all_cat && !wide_seen && max_card <= CELL_MAX_CARD && g->fv_mode == FV_ONE &&
              schedule != DWG_SCHED_SEQUENTIAL && nweight <= CELL_MAX_WEIGHTS && nvar + 64ull * ncolor < CELL_MAX_VARS &&
              !getenv("DWG_NO_CELL");
    g->cell_kmax = (uint32_t)max_card;
  }
  const bool padded = g->sell || g->cell;  // schedule order with classes padded to whole slices

  pt.lap("layout knobs");
  // ---- schedule order: colour-major; SELL graphs additionally sort each class by
  // (is_evidence, degree and others words descending) so that a 32-variable slice is uniform, and pad every
  // class to a multiple of 32 ----
  // SELL: a class is its query variables, padded to a multiple of 32, then its evidence variables,
  // padded again: no slice mixes the two (a query-only sweep skips whole slices, and a learning
  // sweep never meets evidence variables of the first id block behind query variables of the last).
  // A WIDE variable (>= SELL_WIDE_MIN incident records, Boolean sliced-ELL graphs) owns a whole
  // slice: its records are dealt round-robin to the 32 lanes, which sum their shares and combine
  // them by shuffles (SF_WIDE, graph_image.h) -- a hub would otherwise be one serial lane and pad
  // its 31 slice-mates to its length.  Wide variables come first in their part of the class.
  // wide from 32 records on (one full column of the 32 lanes), and only when that is well above the
  // graph's mean: a slice is padded to its longest row, so what hurts is a variable much heavier
  // than the variables around it, not a graph of uniformly heavy ones
  uint64_t nrec_total = 0;
  for (uint64_t v = 0; v < nvar; ++v) nrec_total += inc_end[v] - deg[v];
  uint64_t wide_min = std::max<uint64_t>(SELL_WIDE_MIN, nvar ? 4 * ((nrec_total + nvar - 1) / nvar) : 0);
  if (const char *e = getenv("DWG_WIDE_MIN")) wide_min = (uint64_t)std::max(1, atoi(e));
  std::vector<uint8_t> wide(nvar, 0);
  if (g->sell)
    parallel_for(nvar, [&](uint64_t v0, uint64_t v1) {
      for (uint64_t v = v0; v < v1; ++v) wide[v] = (uint64_t)(inc_end[v] - deg[v]) >= wide_min;
    });
  // slices a wide variable takes: 1, or its number of parts (SELL_PART_RECORDS records each)
  uint64_t part_records = SELL_PART_RECORDS;
  if (const char *e = getenv("DWG_PART_RECORDS")) part_records = (uint64_t)std::max(32, atoi(e));
  auto wide_slices = [&](uint64_t v) -> uint32_t {
    const uint64_t r = inc_end[v] - deg[v];
    return (uint32_t)std::min<uint64_t>(SELL_MAX_PARTS - 1, std::max<uint64_t>(1, (r + part_records - 1) / part_records));
  };
  // per class: [0] query wide, [1] query normal, [2] evidence wide, [3] evidence normal; wslots: slices of the wide ones
  std::vector<std::array<uint32_t, 4>> part(ncolor, std::array<uint32_t, 4>{0, 0, 0, 0});
  std::vector<std::array<uint32_t, 2>> wslots(ncolor, std::array<uint32_t, 2>{0, 0});
  for (uint64_t v = 0; v < nvar; ++v) {
    ++part[color[v]][(g->var_is_evid[v] ? 2 : 0) + (wide[v] ? 0 : 1)];
    if (wide[v]) wslots[color[v]][g->var_is_evid[v] ? 1 : 0] += wide_slices(v);
  }
  auto seg_size = [&](uint32_t c, int ev) { return 32u * wslots[c][ev] + ((part[c][2 * ev + 1] + 31u) & ~31u); };
  g->color_start.assign(ncolor + 1, 0);
  for (uint32_t c = 0; c < ncolor; ++c) {
    uint32_t n = part[c][0] + part[c][1] + part[c][2] + part[c][3];
    if (padded) n = seg_size(c, 0) + seg_size(c, 1);
    g->color_start[c + 1] = g->color_start[c] + n;
  }
  const uint64_t nvar_pad = g->color_start[ncolor];
  g->nvar_pad = nvar_pad;
  g->perm.assign(nvar_pad, 0xffffffffu);
  g->inv.resize(nvar);
  if (padded) {
    // Order inside a class: (is_evidence, id block, then unary / multi / others counts DESCENDING).
    // Evidence last: the evidence tail of a class is one range a query-only sweep never reads.  Id
    // blocks ascending: a LEARNING sweep visits the slices in ascending id block (learn_visit
    // below), because per-sample SGD weights the most recently visited variables most: the
    // reference's results are those of the id-order scan (src/gibbs_sampler.cc:52-54,71-75), and a
    // sweep in another order converges to measurably different weights.  Counts
    // descending inside a block: the 32 variables of a slice have rows of similar length.
    uint64_t id_block = 256;
    while (id_block * 1024 < nvar && id_block < 65536) id_block <<= 1;
    g->id_block = (uint32_t)id_block;
    struct Key { uint64_t hi, lo; uint32_t v; };
    BulkVec<Key> keys(nvar);
    parallel_for(nvar, [&](uint64_t v0, uint64_t v1) {
      for (uint64_t v = v0; v < v1; ++v) {
        uint64_t n_un = 0, n_mu = 0, ow = 0;
        for (uint32_t i = deg[v]; i < inc_end[v]; ++i) {
          uint32_t f = inc[i].fid;
          uint64_t n = raw.fac_ptr[f + 1] - raw.fac_ptr[f];
          int form = finfo[i] & 3;
          if (form == 0) { ++n_un; continue; }
          ++n_mu;
          ow += form == (int)FORM_GENERAL ? 1 + n : (uint64_t)form;
        }
        n_un = std::min<uint64_t>(n_un, (1u << 16) - 1);
        n_mu = std::min<uint64_t>(n_mu, (1u << 12) - 1);
        ow = std::min<uint64_t>(ow, (1u << 20) - 1);
        keys[v] = {((uint64_t)color[v] << 34) | ((uint64_t)g->var_is_evid[v] << 33) | ((uint64_t)(wide[v] ? 0 : 1) << 32) | (v / id_block),
                   ((0xffffull - n_un) << 32) | ((0xfffull - n_mu) << 20) | (0xfffffull - ow), (uint32_t)v};
      }
    });
    // sorted runs in parallel, then pairwise merges (the order is total: ties broken by id)
    auto less = [](const Key &a, const Key &b) {
      return a.hi != b.hi ? a.hi < b.hi : a.lo != b.lo ? a.lo < b.lo : a.v < b.v;
    };
    {
      unsigned runs = 1;
      while (runs * 2 <= host_threads() && nvar / (runs * 2) >= (1u << 16)) runs *= 2;
      auto bound = [&](unsigned r) { return keys.begin() + (std::ptrdiff_t)(nvar * r / runs); };
      parallel_ranges(runs, runs, [&](uint64_t r0, uint64_t r1, unsigned) {
        for (uint64_t r = r0; r < r1; ++r) std::sort(bound((unsigned)r), bound((unsigned)r + 1), less);
      }, 1);
      for (unsigned width = 1; width < runs; width *= 2) {
        const unsigned pairs = runs / (2 * width);
        parallel_ranges(pairs, pairs, [&](uint64_t p0, uint64_t p1, unsigned) {
          for (uint64_t q = p0; q < p1; ++q) {
            const unsigned lo = (unsigned)q * 2 * width;
            std::inplace_merge(bound(lo), bound(lo + width), bound(lo + 2 * width), less);
          }
        }, 1);
      }
    }
    // keys are colour-major: position in the sorted order -> slot of the (padded) class
    std::vector<uint64_t> class_first(ncolor + 1, 0);
    for (uint32_t c = 0; c < ncolor; ++c)
      class_first[c + 1] = class_first[c] + part[c][0] + part[c][1] + part[c][2] + part[c][3];
    parallel_for(nvar, [&](uint64_t i0, uint64_t i1) {
      for (uint64_t i = i0; i < i1; ++i) {
        const uint32_t v = keys[i].v, c = color[v];
        uint64_t k = i - class_first[c], base = g->color_start[c];
        const int ev = k >= (uint64_t)part[c][0] + part[c][1];
        if (ev) { k -= (uint64_t)part[c][0] + part[c][1]; base += seg_size(c, 0); }
        // inside the part: wide variables (one slice, or one per part of a hub), then the others packed
        const uint64_t nw = part[c][2 * ev];
        if (k < nw) continue;  // placed below: their slots depend on the parts of the wide variables before them
        const uint32_t d = (uint32_t)(base + 32ull * wslots[c][ev] + (k - nw));
        g->perm[d] = v;
        g->inv[v] = d;
      }
    });
    for (uint32_t c = 0; c < ncolor; ++c)
      for (int ev = 0; ev < 2; ++ev) {
        uint64_t i = class_first[c] + (ev ? (uint64_t)part[c][0] + part[c][1] : 0);
        uint64_t d = g->color_start[c] + (ev ? seg_size(c, 0) : 0);
        for (uint32_t k = 0; k < part[c][2 * ev]; ++k, ++i) {
          const uint32_t v = keys[i].v;
          g->perm[d] = v;
          g->inv[v] = (uint32_t)d;
          d += 32ull * wide_slices(v);
        }
      }

  } else {
    std::vector<uint32_t> fill(g->color_start.begin(), g->color_start.end() - 1);
    for (uint64_t v = 0; v < nvar; ++v) {
      uint32_t d = fill[color[v]]++;
      g->perm[d] = (uint32_t)v;
      g->inv[v] = d;
    }
  }

  pt.lap("schedule order (sort)");
  // ---- emit rows and edge records (CSR) in schedule order ----
  const uint64_t ndrow = g->sell ? nvar_pad : nrow + (nvar_pad - nvar);  // device rows (pads own an empty row)
  g->ndrow = ndrow;
  g->var_flags.assign(nvar_pad, VF_PAD); g->evid_val.assign(nvar_pad, 0);
  if (g->has_categorical) g->var_row.resize(nvar_pad + 1);
  g->row_ptr.assign(ndrow + 1, 0);
  if (g->has_multi) g->row_oth.assign(ndrow + 1, 0);
  if (g->has_truthiness) g->truthiness.assign(ndrow, 0.0);
  g->row_of_value.resize(nrow);
  uint64_t nindex = 0;
  for (uint64_t v = 0; v < nvar; ++v) nindex += inc_end[v] - deg[v];
  if (nindex >= U32_LIMIT) { *err = DWG_ERR_UNSUPPORTED; return "more than 2^32 edge records"; }
  g->nindex = nindex;
  g->e_wid.resize(nindex); g->e_tag.resize(nindex);
  if (g->fv_mode == FV_F32) g->e_fv32.resize(nindex);
  if (g->fv_mode == FV_F64) g->e_fv64.resize(nindex);
  const int ow = g->oth_packed ? 1 : 2;  // words per others entry
  // Two passes over contiguous blocks of schedule positions, both in parallel: pass 1 sizes each
  // block (rows, edge records, others entries), a prefix sum places it, pass 2 writes it.
  struct Block {
    uint64_t rows = 0, edges = 0, oth = 0;   // sizes, then (after the prefix sum) start offsets
    uint64_t bad_d = INVALID;
    const char *msg = nullptr;
    double b_all = 0, b_query = 0, b_learn = 0;
  };
  const unsigned NB = nvar_pad >= (1u << 12) ? host_threads() * 4 : 1;
  std::vector<Block> blk(NB + 1);
  parallel_ranges(nvar_pad, NB, [&](uint64_t d0, uint64_t d1, unsigned p) {
    Block &B = blk[p];
    for (uint64_t d = d0; d < d1; ++d) {
      uint64_t v = g->perm[d];
      if (v == 0xffffffffu) { ++B.rows; continue; }  // pad: an empty row
      bool cat = g->var_dtype[v];
      B.rows += cat ? raw.var_card[v] : 1;
      if (adense[v] >= U32_LIMIT && B.bad_d == INVALID) { B.bad_d = d; B.msg = "initial value >= 2^32"; }
      for (uint32_t i = deg[v]; i < inc_end[v]; ++i) {
        uint32_t f = inc[i].fid;
        uint64_t n = raw.fac_ptr[f + 1] - raw.fac_ptr[f];
        int form = finfo[i] & 3;
        if (form == (int)FORM_GENERAL && n >= (g->oth_packed ? (1ull << 31) : U32_LIMIT) && B.bad_d == INVALID) {
          B.bad_d = d; B.msg = "factor arity too large";
        }
        B.oth += form != (int)FORM_GENERAL ? (uint64_t)form : 1 + n;
        ++B.edges;
      }
    }
  });
  {
    uint64_t rows = 0, edges = 0, oth = 0;
    for (unsigned p = 0; p <= NB; ++p) {
      if (p < NB && blk[p].bad_d != INVALID) { *err = DWG_ERR_UNSUPPORTED; return blk[p].msg; }
      uint64_t r = blk[p].rows, e = blk[p].edges, o = blk[p].oth;
      blk[p].rows = rows; blk[p].edges = edges; blk[p].oth = oth;
      rows += r; edges += e; oth += o;
    }
    if (blk[NB].oth >= U32_LIMIT) { *err = DWG_ERR_UNSUPPORTED; return "others stream >= 2^32 entries"; }
    g->oth.assign(blk[NB].oth * ow, 0);
  }
  parallel_ranges(nvar_pad, NB, [&](uint64_t d0, uint64_t d1, unsigned p) {
    Block &B = blk[p];
    uint64_t row = B.rows, ei = B.edges, oc = B.oth;
    auto put_oth = [&](uint32_t a, uint32_t eq) {
      if (g->oth_packed) g->oth[oc] = (a << 1) | eq;
      else { g->oth[2 * oc] = a; g->oth[2 * oc + 1] = eq; }
      ++oc;
    };
    for (uint64_t d = d0; d < d1; ++d) {
      uint64_t v = g->perm[d];
      if (v == 0xffffffffu) {  // pad: an empty row
        if (g->has_categorical) g->var_row[d] = (uint32_t)row;
        g->row_ptr[row] = (uint32_t)ei;
        if (g->has_multi) g->row_oth[row] = (uint32_t)oc;
        ++row;
        continue;
      }
      bool cat = g->var_dtype[v];
      uint64_t K = cat ? raw.var_card[v] : 1;
      g->var_flags[d] = (g->var_is_evid[v] ? VF_EVID : 0) | (cat ? VF_CATEGORICAL : 0) |
                        (has_truth[v] ? VF_TRUTHINESS : 0);
      g->evid_val[d] = (uint32_t)adense[v];
      if (g->has_categorical) g->var_row[d] = (uint32_t)row;
      double vb = 13.0 + (cat ? 4.0 * K : 0.0), vl = 2.0;
      uint32_t i = deg[v];
      for (uint64_t k = 0; k < K; ++k, ++row) {
        g->row_of_value[g->value_base[v] + k] = (uint32_t)row;
        if (g->has_truthiness) g->truthiness[row] = truth_orig[g->value_base[v] + k];
        g->row_ptr[row] = (uint32_t)ei;
        if (g->has_multi) g->row_oth[row] = (uint32_t)oc;
        for (; i < inc_end[v] && inc[i].key == k; ++i, ++ei) {
          uint32_t f = inc[i].fid;
          uint64_t a = raw.fac_ptr[f], b = raw.fac_ptr[f + 1], n = b - a;
          g->e_wid[ei] = (uint32_t)raw.fac_wid[f];
          if (g->fv_mode == FV_F32) g->e_fv32[ei] = (float)raw.fac_fv[f];
          if (g->fv_mode == FV_F64) g->e_fv64[ei] = raw.fac_fv[f];
          uint8_t tag = (uint8_t)func_code(raw.fac_func[f]);
          const uint64_t self_pos = finfo[i] >> 2;
          int form = finfo[i] & 3;
          if (form != (int)FORM_GENERAL) {
            if (!cat && edense[a + self_pos] == 1) tag |= TAG_SELF_EQ1;
            if (self_pos == n - 1) tag |= TAG_SELF_HEAD;
            tag |= (uint8_t)(form << TAG_FORM_SHIFT);
            for (uint64_t e = a; e < b; ++e)
              if (raw.edge_vid[e] != v) put_oth(g->inv[raw.edge_vid[e]], edense[e]);
          } else {
            // general form: length word, then every position (self included) in order
            tag |= (uint8_t)(FORM_GENERAL << TAG_FORM_SHIFT);
            put_oth((uint32_t)n >> (g->oth_packed ? 1 : 0), g->oth_packed ? (uint32_t)n & 1u : 0u);  // the word n
            for (uint64_t e = a; e < b; ++e) put_oth(g->inv[raw.edge_vid[e]], edense[e]);
          }
          g->e_tag[ei] = tag;
          vb += 9.0 + 5.0 * (double)(n - 1);
          vl += 1.0 * (double)(n - 1);
        }
      }
      B.b_all += vb;
      if (!g->var_is_evid[v]) B.b_query += vb;
      B.b_learn += vb + vl;
    }
  });
  const uint64_t ei = blk[NB].edges;
  double b_all = 0, b_query = 0, b_learn = 0;
  for (unsigned p = 0; p < NB; ++p) { b_all += blk[p].b_all; b_query += blk[p].b_query; b_learn += blk[p].b_learn; }
  g->row_ptr[ndrow] = (uint32_t)ei;
  if (g->has_multi) g->row_oth[ndrow] = (uint32_t)(g->oth.size() / ow);
  if (g->has_categorical) g->var_row[nvar_pad] = (uint32_t)ndrow;
  g->noth = g->oth.size() / ow;
  g->bytes_infer = b_all;
  g->bytes_infer_query = b_query;
  g->bytes_learn = b_learn;

  if (g->assign_bytes == 1) {
    g->evid_val8.resize(nvar_pad);
    parallel_for(nvar_pad, [&](uint64_t a, uint64_t b) {
      for (uint64_t d = a; d < b; ++d) g->evid_val8[d] = (uint8_t)g->evid_val[d];
    });
  }
  pt.lap("emit CSR");
  // ---- internal weight ids: hottest first, so that ids < n_hot can live in shared memory ----
  g->w_perm.resize(nweight); g->w_inv.resize(nweight);
  {
    std::iota(g->w_perm.begin(), g->w_perm.end(), 0u);
    if (nweight > DWG_HOT_WEIGHTS) {
      // use counts: one histogram per thread, summed (integers: any order gives the same counts)
      const unsigned T = host_threads();
      std::vector<std::vector<uint32_t>> part(T);
      parallel_ranges(nindex, T, [&](uint64_t e0, uint64_t e1, unsigned p) {
        part[p].assign(nweight, 0);
        for (uint64_t e = e0; e < e1; ++e) ++part[p][g->e_wid[e]];
      });
      std::vector<uint64_t> cnt(nweight, 0);
      for (unsigned p = 0; p < T; ++p)
        if (!part[p].empty())
          for (uint64_t w = 0; w < nweight; ++w) cnt[w] += part[p][w];
      std::stable_sort(g->w_perm.begin(), g->w_perm.end(),
                       [&](uint32_t a, uint32_t b) { return cnt[a] > cnt[b]; });
    }
    for (uint64_t i = 0; i < nweight; ++i) g->w_inv[g->w_perm[i]] = (uint32_t)i;
    if (nweight > DWG_HOT_WEIGHTS)
      parallel_for(nindex, [&](uint64_t e0, uint64_t e1) {
        for (uint64_t e = e0; e < e1; ++e) g->e_wid[e] = g->w_inv[g->e_wid[e]];
      });
    g->n_hot = (uint32_t)std::min<uint64_t>(nweight, DWG_HOT_WEIGHTS);
    if (const char *e = getenv("DWG_HOT_WEIGHTS"))  // tests: a small shared-memory copy, most weights read from / pushed to L2
      g->n_hot = (uint32_t)std::min<uint64_t>(g->n_hot, (uint64_t)std::max(16, atoi(e)));
  }

  pt.lap("weight ids");
  class_statistics(raw, g, nvar_pad, ncolor, nweight, ow);
  pt.lap("class statistics");
  {
    std::string e = build_sell_blob(raw, g, nvar, nvar_pad, ncolor, wide, wide_slices, err);
    if (!e.empty()) return e;
  }
  {
    std::string e = build_cell_blob(g, nvar_pad, ncolor, err);
    if (!e.empty()) return e;
  }
  pt.lap("sliced-ELL blob");
  *err = DWG_OK;
  return "";
}

}  // namespace dwg
