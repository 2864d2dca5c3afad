// synth.cc -- streaming builder of the device image of a SYNTHETIC unary Boolean graph
// (BASELINE.json configs[3] "logistic 50M vars x 20 features, 1M tied weights" and configs[4]
// "~300M vars x 10 features per replica"): 10^9 .. 3 x 10^9 factors, far beyond what the decoded
// records (RawGraph, ~50 B per factor) or the reference's own FactorGraph (72 B per edge) can hold
// in host memory.  Every record is a pure function of (seed, stream, index) -- the splitmix64
// finaliser, mirrored in numpy by sampler_b200/graphgen.py synth_unary -- so the builder writes
// the sliced-ELL blob slice by slice without ever materialising the factors.
//
// The image is what compile_graph (compile.cc) makes of the same records: same schedule order
// (query variables, then evidence variables, ascending id), same record words, same weight
// statistics, same learning order; tests/test_synth.py compares the two array by array at sizes
// the record path can hold.  Only the CSR stream (parity hooks, sequential schedule) is not built.
// Input tooling for the benchmark, not part of the reference's interface.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../../include/dwg.h"
#include "../graph_image.h"
#include "parallel.h"

namespace dwg {
namespace {

inline uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
constexpr double SQRT3 = 1.7320508075688772;

struct Gen {
  uint64_t base[12];
  uint64_t nweight;
  bool real_features;
  const double *cdf = nullptr;       // Zipf table or null (uniform ids)
  std::vector<uint32_t> guide;       // guide[g] = first index with cdf >= g / G
  static constexpr uint32_t G = 1u << 16;
  explicit Gen(uint64_t seed) {
    for (uint64_t s = 0; s < 12; ++s) base[s] = mix64(seed + s * 0xD1B54A32D192ED03ull);
  }
  uint64_t h(int stream, uint64_t i) const { return mix64(base[stream] + i); }
  uint64_t wid(uint64_t f) const {
    const uint64_t x = h(3, f);
    if (!cdf) return x % nweight;
    const double u = u01(x);
    const uint32_t g = std::min<uint32_t>((uint32_t)(u * G), G - 1);
    const double *lo = cdf + guide[g], *hi = cdf + std::min<uint64_t>(nweight, (uint64_t)guide[g + 1] + 1);
    const uint64_t idx = (uint64_t)(std::lower_bound(lo, hi, u) - cdf);  // first index with cdf[idx] >= u
    return std::min(idx, nweight - 1);
  }
  double fv(uint64_t f) const {
    if (!real_features) return 1.0;
    return (((u01(h(4, f)) + u01(h(5, f))) + u01(h(6, f))) + u01(h(7, f)) - 2.0) * SQRT3;
  }
  double wstar(uint64_t k) const {
    return 0.5 * ((((u01(h(8, k)) + u01(h(9, k))) + u01(h(10, k))) + u01(h(11, k)) - 2.0) * SQRT3);
  }
};

}  // namespace

std::string build_synthetic_unary(uint64_t nvar, uint32_t degree, uint64_t nweight, double zipf, bool real_features,
                                  double p_evid, uint64_t seed, GraphImage *g, std::vector<double> *cdf_out, int *err) {
  *err = DWG_ERR_INVALID;
  *g = GraphImage();
  if (!nvar || !degree || !nweight) return "synthetic graph: empty";
  if (nvar >= 0x7fffffffull || nweight > SELL_MAX_WEIGHTS || nvar * (uint64_t)degree >= 0xffffffffull * 4) {
    *err = DWG_ERR_UNSUPPORTED;
    return "synthetic graph exceeds the 32-bit device layout";
  }
  const uint64_t nfac = nvar * degree;
  Gen gen(seed);
  gen.nweight = nweight;
  gen.real_features = real_features;
  std::vector<double> cdf;
  if (zipf > 0) {
    cdf.resize(nweight);
    double acc = 0;
    for (uint64_t r = 0; r < nweight; ++r) { acc += std::pow((double)(r + 1), -zipf); cdf[r] = acc; }
    for (uint64_t r = 0; r < nweight; ++r) cdf[r] /= acc;
    gen.cdf = cdf.data();
    gen.guide.resize(Gen::G + 1);
    for (uint32_t q = 0; q <= Gen::G; ++q)
      gen.guide[q] = (uint32_t)std::min<uint64_t>(nweight - 1, std::lower_bound(cdf.begin(), cdf.end(), (double)q / Gen::G) - cdf.begin());
    if (cdf_out) *cdf_out = cdf;
  }
  std::vector<double> wstar(nweight);
  parallel_for(nweight, [&](uint64_t a, uint64_t b) { for (uint64_t k = a; k < b; ++k) wstar[k] = gen.wstar(k); });

  // ---- sizes and flags (compile.cc "variables", "layout knobs") ----
  g->nvar = nvar; g->nfac = nfac; g->nweight = nweight; g->nedge = nfac; g->nrow = nvar; g->nindex = nfac;
  g->schedule = DWG_SCHED_HOGWILD;
  g->has_categorical = false; g->has_multi = false; g->oth_packed = true;
  g->fv_mode = real_features ? FV_F64 : FV_ONE;
  g->assign_bytes = 1;
  g->sell = true;
  g->csr_built = false;
  g->w_init.assign(nweight, 0.0);
  g->w_fixed.assign(nweight, 0);
  g->var_dtype.assign(nvar, 0);
  g->var_is_evid.resize(nvar);
  const unsigned T = host_threads();
  std::vector<uint64_t> part_q(T + 1, 0);
  parallel_ranges(nvar, T, [&](uint64_t a, uint64_t b, unsigned p) {
    uint64_t q = 0;
    for (uint64_t v = a; v < b; ++v) {
      const bool ev = u01(gen.h(1, v)) < p_evid;
      g->var_is_evid[v] = ev;
      q += !ev;
    }
    part_q[p + 1] = q;
  }, 1);
  for (unsigned p = 0; p < T; ++p) part_q[p + 1] += part_q[p];
  const uint64_t nq = part_q[T], ne = nvar - nq;
  g->nvar_query = nq; g->nvar_evid = ne;
  const uint64_t qpad = (nq + 31) & ~31ull, nvar_pad = qpad + ((ne + 31) & ~31ull);
  g->nvar_pad = nvar_pad; g->ndrow = nvar_pad;
  g->color_start = {0u, (uint32_t)nvar_pad};
  uint64_t id_block = 256;
  while (id_block * 1024 < nvar && id_block < 65536) id_block <<= 1;
  g->id_block = (uint32_t)id_block;

  // ---- schedule order: query variables in ascending id, padding, evidence variables in ascending id ----
  g->perm.assign(nvar_pad, 0xffffffffu);
  g->inv.resize(nvar);
  g->var_flags.assign(nvar_pad, VF_PAD);
  g->evid_val8.assign(nvar_pad, 0);
  parallel_ranges(nvar, T, [&](uint64_t a, uint64_t b, unsigned p) {
    uint64_t q = part_q[p], e = a - part_q[p];  // query / evidence variables before `a`
    for (uint64_t v = a; v < b; ++v) {
      const bool ev = g->var_is_evid[v];
      const uint64_t d = ev ? qpad + e++ : q++;
      g->perm[d] = (uint32_t)v;
      g->inv[v] = (uint32_t)d;
      g->var_flags[d] = ev ? VF_EVID : 0;
    }
  }, 1);
  g->row_of_value = g->inv;  // a Boolean variable owns one tally row: its schedule position

  // ---- internal weight ids: hottest first (compile.cc "weight ids") ----
  g->w_perm.resize(nweight); g->w_inv.resize(nweight);
  std::iota(g->w_perm.begin(), g->w_perm.end(), 0u);
  if (nweight > DWG_HOT_WEIGHTS) {
    std::vector<std::vector<uint32_t>> part(T);
    parallel_ranges(nfac, T, [&](uint64_t f0, uint64_t f1, unsigned p) {
      part[p].assign(nweight, 0);
      for (uint64_t f = f0; f < f1; ++f) ++part[p][gen.wid(f)];
    }, 1);
    std::vector<uint64_t> cnt(nweight, 0);
    for (unsigned p = 0; p < T; ++p)
      for (uint64_t w = 0; w < nweight; ++w) cnt[w] += part[p][w];
    std::stable_sort(g->w_perm.begin(), g->w_perm.end(), [&](uint32_t a, uint32_t b) { return cnt[a] > cnt[b]; });
  }
  for (uint64_t i = 0; i < nweight; ++i) g->w_inv[g->w_perm[i]] = (uint32_t)i;
  g->n_hot = (uint32_t)std::min<uint64_t>(nweight, DWG_HOT_WEIGHTS);

  // ---- sliced-ELL blob + per-weight learning statistics, slice by slice ----
  const uint64_t nslice = nvar_pad / 32, fvw = real_features ? 2 : 0;
  const uint64_t slice_words = SH_WORDS + 32ull * degree * (1 + fvw);
  if (nslice * slice_words / 8 >= 0xffffffffull) { *err = DWG_ERR_UNSUPPORTED; return "sliced-ELL stream exceeds 2^32 x 32 B"; }
  g->sell_max_words = (uint32_t)slice_words;
  g->sl_off.resize(nslice + 1);
  parallel_for(nslice + 1, [&](uint64_t a, uint64_t b) { for (uint64_t s = a; s < b; ++s) g->sl_off[s] = (uint32_t)(s * slice_words / 8); });
  g->blob.resize(nslice * slice_words);  // every word is written below (headers, records, padding)
  // statistics as 2^-20 fixed point (compile.cc): hot weights per thread, the rest with atomics
  const uint64_t n_hot = g->n_hot;
  std::vector<uint64_t> cold(6 * nweight, 0);
  std::vector<std::vector<uint64_t>> hot(T, std::vector<uint64_t>(6 * n_hot, 0));
  auto fx = [](double x) { return (uint64_t)std::llround(x * 1048576.0); };
  parallel_ranges(nslice, T, [&](uint64_t s0, uint64_t s1, unsigned p) {
    std::vector<uint32_t> wids(degree);
    std::vector<double> fvs(degree);
    uint64_t *myhot = hot[p].data();
    for (uint64_t s = s0; s < s1; ++s) {
      uint32_t *h = g->blob.data() + s * slice_words;
      std::memset(h, 0, slice_words * 4);
      h[SH_N_UNARY] = degree; h[SH_SELF_OFF] = g->sl_off[s]; h[SH_FLAGS] = SF_ANTISYM;
      uint32_t *un = h + SH_WORDS, *unfv = un + 32ull * degree;
      for (uint32_t l = 0; l < 32; ++l) {
        const uint64_t d = s * 32 + l;
        const uint64_t v = g->perm[d];
        if (v == 0xffffffffu) {
          h[SH_PAD_MASK] |= 1u << l;
          for (uint32_t j = 0; j < degree; ++j) un[32ull * j + l] = SELL_UNARY_NOP;
          continue;
        }
        const bool ev = g->var_is_evid[v];
        double acc = 0, rowsum = 0;
        for (uint32_t j = 0; j < degree; ++j) {
          const uint64_t f = v * degree + j;
          const uint64_t w = gen.wid(f);
          const double x = gen.fv(f);
          wids[j] = g->w_inv[w]; fvs[j] = x;
          acc = acc + wstar[w] * x;
          rowsum += std::fabs(x) * 2.0;
          un[32ull * j + l] = sell_unary_record(wids[j], false, 2u, 0u);  // ISTRUE(x = 1): +1 when 1, -1 when 0
          if (fvw) std::memcpy(unfv + (32ull * j + l) * 2, &x, 8);
        }
        // planted logistic label; evidence variables keep it, query variables start at 0 (binary_format.cc:107-111)
        const bool label = u01(gen.h(2, v)) < 1.0 / (1.0 + std::exp(-(2.0 * acc)));
        if (ev) {
          h[SH_EVID_MASK] |= 1u << l;
          if (label) { h[SH_EVVAL_MASK] |= 1u << l; g->evid_val8[d] = 1; }
        }
        for (uint32_t j = 0; j < degree; ++j) {
          const double delta = std::fabs(fvs[j]) * 2.0;
          const uint64_t q = fx(0.25 * delta * rowsum), hd = fx(0.25 * delta * delta);
          const uint32_t w = wids[j];
          if (w < n_hot) {
            uint64_t *a = myhot + 6ull * w;
            a[0] += 1ull << 20; a[1] += q; a[2] += hd;
            if (ev) { a[3] += 1ull << 20; a[4] += q; a[5] += hd; }
          } else {
            uint64_t *a = cold.data() + 6ull * w;
            __atomic_fetch_add(&a[0], 1ull << 20, __ATOMIC_RELAXED);
            __atomic_fetch_add(&a[1], q, __ATOMIC_RELAXED);
            __atomic_fetch_add(&a[2], hd, __ATOMIC_RELAXED);
            if (ev) {
              __atomic_fetch_add(&a[3], 1ull << 20, __ATOMIC_RELAXED);
              __atomic_fetch_add(&a[4], q, __ATOMIC_RELAXED);
              __atomic_fetch_add(&a[5], hd, __ATOMIC_RELAXED);
            }
          }
        }
      }
    }
  }, 1);
  g->cls_shared = true;
  g->cls_N_all.resize(nweight); g->cls_Q_all.resize(nweight); g->cls_H_all.resize(nweight);
  g->cls_N_evid.resize(nweight); g->cls_Q_evid.resize(nweight); g->cls_H_evid.resize(nweight);
  parallel_for(nweight, [&](uint64_t a, uint64_t b) {
    for (uint64_t w = a; w < b; ++w) {
      uint64_t t[6];
      for (int i = 0; i < 6; ++i) t[i] = cold[6 * w + i];
      if (w < n_hot)
        for (unsigned p = 0; p < T; ++p)
          for (int i = 0; i < 6; ++i) t[i] += hot[p][6 * w + i];
      g->cls_N_all[w] = (float)((double)t[0] / 1048576.0); g->cls_Q_all[w] = (float)((double)t[1] / 1048576.0);
      g->cls_H_all[w] = (float)((double)t[2] / 1048576.0); g->cls_N_evid[w] = (float)((double)t[3] / 1048576.0);
      g->cls_Q_evid[w] = (float)((double)t[4] / 1048576.0); g->cls_H_evid[w] = (float)((double)t[5] / 1048576.0);
    }
  });

  // ---- learning order (compile.cc): slices by id block of their first variable ----
  {
    std::vector<uint64_t> vk(nslice);
    parallel_for(nslice, [&](uint64_t a, uint64_t b) {
      for (uint64_t s = a; s < b; ++s) {
        uint64_t v = 0;
        for (uint64_t d = s * 32; d < s * 32 + 32; ++d)
          if (g->perm[d] != 0xffffffffu) { v = g->perm[d]; break; }
        vk[s] = v / id_block;
      }
    });
    g->learn_visit.resize(nslice);
    std::iota(g->learn_visit.begin(), g->learn_visit.end(), 0u);
    // two ascending runs (query slices, evidence slices): a stable merge by block
    const uint32_t mid = (uint32_t)(qpad / 32);
    std::inplace_merge(g->learn_visit.begin(), g->learn_visit.begin() + mid, g->learn_visit.end(),
                       [&](uint32_t a, uint32_t b) { return vk[a] != vk[b] ? vk[a] < vk[b] : a < b; });
  }

  // ---- ALGORITHMIC bytes per sweep: 13 per variable + 9 per unary record; learning + 2 ----
  const double vb = 13.0 + 9.0 * degree;
  g->bytes_infer = vb * (double)nvar;
  g->bytes_infer_query = vb * (double)nq;
  g->bytes_learn = (vb + 2.0) * (double)nvar;
  *err = DWG_OK;
  return "";
}

}  // namespace dwg
