/* dw_oracle.c -- CPU ORACLE for the DimmWitted Gibbs hot path (see dw_oracle.h).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Parity PINNED against the reference's
 * own known-answer tests and against the reference objects themselves
 * (tests/test_oracle_*.py).  All citations are /root/reference/<file>:<line>.
 * Compile with -ffp-contract=off: the reference is built for baseline x86-64
 * (no FMA), so every multiply-add below rounds twice, as it does there.
 */
#define _DEFAULT_SOURCE
#define _XOPEN_SOURCE 600
#include "dw_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* src/common.h:13-15 */
#define LOG_2 0.693147180559945
#define MINUS_LOG_THRESHOLD -18.42
#define LINEAR_ZERO_THRESHOLD 0.000001

/* src/common.h:111-113 */
static int is_linear_zero(double x) {
  return x <= LINEAR_ZERO_THRESHOLD && x >= -LINEAR_ZERO_THRESHOLD;
}

/* src/common.h:118-132 */
double orc_logadd(double log_a, double log_b) {
  if (log_a < log_b) {
    double t = log_a;
    log_a = log_b;
    log_b = t;
  } else if (log_a <= log_b && log_b <= log_a) {
    return LOG_2 + log_a;
  }
  double nad = log_b - log_a;
  if (nad < MINUS_LOG_THRESHOLD) return log_a;
  return log_a + log1p(exp(nad));
}

/* src/factor.h:94-100 + src/variable.h:160-162 */
static int sat(uint64_t vif_vid, uint64_t vif_eq, uint64_t vid,
               const uint64_t *assignments, uint64_t proposal) {
  return (vif_vid == vid) ? (vif_eq == proposal) : (vif_eq == assignments[vif_vid]);
}

/* Factor::potential, src/factor.h:59-86 and the POTENTIAL_SIGN bodies :112-299 */
double orc_factor_potential(int func, uint64_t n, const uint64_t *vids,
                            const uint64_t *eqs, double fv,
                            const uint64_t *a, uint64_t vid, uint64_t p) {
  double sign;
  uint64_t i;
#define SAT(i) sat(vids[i], eqs[i], vid, a, p)
  switch (func) {
    case ORC_FUNC_EQUAL: { /* :112-128 */
      int first = SAT(0);
      sign = 1;
      for (i = 0; i < n; ++i)
        if (SAT(i) != first) {
          sign = -1;
          break;
        }
      break;
    }
    case ORC_FUNC_AND:
    case ORC_FUNC_ISTRUE: { /* :135-145, alias :73 */
      sign = 1;
      for (i = 0; i < n; ++i)
        if (!SAT(i)) {
          sign = -1;
          break;
        }
      break;
    }
    case ORC_FUNC_AND_CATEGORICAL: { /* :148-158 */
      sign = 1;
      for (i = 0; i < n; ++i)
        if (!SAT(i)) {
          sign = 0;
          break;
        }
      break;
    }
    case ORC_FUNC_OR: { /* :165-175 */
      sign = -1;
      for (i = 0; i < n; ++i)
        if (SAT(i)) {
          sign = 1;
          break;
        }
      break;
    }
    case ORC_FUNC_IMPLY_MLN: { /* :189-209 */
      int body = 1;
      for (i = 0; i + 1 < n; ++i) body &= SAT(i);
      sign = !body ? 1 : (SAT(n - 1) ? 1 : 0);
      break;
    }
    case ORC_FUNC_IMPLY_NATURAL: { /* :223-243 */
      int body = 1;
      for (i = 0; i + 1 < n; ++i) body &= SAT(i);
      sign = !body ? 0 : (SAT(n - 1) ? 1 : -1);
      break;
    }
    case ORC_FUNC_LINEAR: { /* :246-261 */
      double res = 0.0;
      int head = SAT(n - 1);
      for (i = 0; i + 1 < n; ++i) res += ((1 - SAT(i)) || head);
      sign = (n == 1) ? (double)head : res;
      break;
    }
    case ORC_FUNC_RATIO: { /* :264-277 */
      double res = 1.0;
      int head = SAT(n - 1);
      for (i = 0; i + 1 < n; ++i) res += ((1 - SAT(i)) || head);
      sign = (n == 1) ? log2(res + (double)head) : log2(res);
      break;
    }
    case ORC_FUNC_LOGICAL: { /* :280-299 */
      double res = 0.0;
      int head = SAT(n - 1);
      for (i = 0; i + 1 < n; ++i) res += ((1 - SAT(i)) || head);
      sign = (n == 1) ? (double)head : (res > 0.0 ? 1.0 : 0.0);
      break;
    }
    default: /* :81-84 aborts */
      abort();
  }
#undef SAT
  return sign * fv; /* :65-66 */
}

static double factor_pot(const orc_graph *g, uint64_t fid, const uint64_t *a,
                         uint64_t vid, uint64_t p) {
  uint64_t b = g->vif_base[fid];
  return orc_factor_potential(g->func[fid], g->num_vars[fid], g->vif_vid + b,
                              g->vif_eq + b, g->fv[fid], a, vid, p);
}

/* FactorGraph::potential, src/factor_graph.h:127-145 */
double orc_potential(const orc_graph *g, uint64_t vid, uint64_t proposal,
                     const uint64_t *a, const double *weights) {
  double pot = 0.0;
  uint64_t off = g->dtype[vid] == 0 ? 0 : proposal; /* var_value_offset, variable.h:113 */
  uint64_t row = g->var_val_base[vid] + off;
  for (uint64_t i = 0; i < g->fi_len[row]; ++i) {
    uint64_t fid = g->factor_index[g->fi_base[row] + i];
    double w = weights[g->wid[fid]];
    pot += w * factor_pot(g, fid, a, vid, proposal);
  }
  return pot;
}

/* src/gibbs_sampler.h:198-215 */
uint64_t orc_draw_boolean(double pot_pos, double pot_neg, double r) {
  return (r * (1.0 + exp(pot_neg - pot_pos)) < 1.0) ? 1 : 0;
}

/* src/gibbs_sampler.h:217-246 (pots already evaluated; sum by sequential logadd) */
uint64_t orc_draw_categorical(const double *pots, uint64_t card, double r) {
  double sum = -100000.0;
  for (uint64_t i = 0; i < card; ++i) sum = orc_logadd(sum, pots[i]);
  for (uint64_t i = 0; i < card; ++i) {
    r -= exp(pots[i] - sum);
    if (r <= 0) return i;
  }
  return ORC_INVALID; /* reference: assert fails :244 */
}

/* GibbsSamplerThread::draw_sample with the uniform made explicit */
uint64_t orc_draw_sample_r(const orc_graph *g, uint64_t vid, const uint64_t *a,
                           const double *w, double r) {
  if (g->dtype[vid] == 0) {
    double pos = orc_potential(g, vid, 1, a, w);
    double neg = orc_potential(g, vid, 0, a, w);
    return orc_draw_boolean(pos, neg, r);
  }
  uint64_t card = g->card[vid];
  double *buf = (double *)malloc(sizeof(double) * (card ? card : 1));
  for (uint64_t i = 0; i < card; ++i) buf[i] = orc_potential(g, vid, i, a, w);
  uint64_t p = orc_draw_categorical(buf, card, r);
  free(buf);
  return p;
}

/* InferenceResult::update_weight, src/inference_result.h:66-85 */
void orc_update_weight(const orc_opts *o, double *weights, uint64_t wid,
                       double stepsize, double gradient) {
  double diff = stepsize * gradient;
  double w = weights[wid];
  if (!o->reg_l1)
    w *= (1.0 / (1.0 + o->reg_param * stepsize));
  else
    w += o->reg_param * (w < 0);
  w -= diff;
  weights[wid] = w;
}

/* FactorGraph::sgd_on_factor, src/factor_graph.cc:243-260 */
static void sgd_on_factor(const orc_graph *g, const orc_opts *o, orc_state *s,
                          uint64_t fid, double stepsize, uint64_t vid,
                          uint64_t evidence_value) {
  if (g->w_fixed[g->wid[fid]]) return;
  double pot_evid = factor_pot(g, fid, s->assign_evid, vid, evidence_value);
  double pot_free = factor_pot(g, fid, s->assign_free, ORC_INVALID, ORC_INVALID);
  orc_update_weight(o, s->weights, g->wid[fid], stepsize, pot_free - pot_evid);
}

/* FactorGraph::sgd_on_variable, src/factor_graph.cc:262-314 */
void orc_sgd_on_variable(const orc_graph *g, const orc_opts *o, orc_state *s,
                         uint64_t vid, double stepsize) {
  if (g->dtype[vid] == 0) {
    uint64_t row = g->var_val_base[vid];
    for (uint64_t j = 0; j < g->fi_len[row]; ++j)
      sgd_on_factor(g, o, s, g->factor_index[g->fi_base[row] + j], stepsize, vid,
                    g->assignment_dense[vid]);
    return;
  }
  uint64_t proposal = s->assign_free[vid];
  for (uint64_t val = 0; val < g->card[vid]; ++val) {
    if (!o->noise_aware && val != g->assignment_dense[vid]) continue;
    uint64_t ev = g->var_val_base[vid] + val;
    if (o->noise_aware && is_linear_zero(g->truthiness[ev])) continue;
    double t = o->noise_aware ? g->truthiness[ev] : 1;
    for (uint64_t j = 0; j < g->fi_len[ev]; ++j)
      sgd_on_factor(g, o, s, g->factor_index[g->fi_base[ev] + j], stepsize * t, vid, val);
    if (val == proposal) continue;
    uint64_t pv = g->var_val_base[vid] + proposal;
    for (uint64_t j = 0; j < g->fi_len[pv]; ++j)
      sgd_on_factor(g, o, s, g->factor_index[g->fi_base[pv] + j], stepsize * t, vid, val);
  }
}

/* Uniform source: the reference consumes one erand48 stream per thread
 * (src/gibbs_sampler.h:203,237,177).  For schedule-order parity tests the same code can be fed
 * explicit per-variable uniforms instead: u[0] for the first draw of the step, u[1] for the
 * second. */
typedef struct {
  unsigned short *seed; /* erand48 state, or NULL */
  double u[2];
  int k;
} usrc;
static double next_u(usrc *s) {
  if (s->seed) return erand48(s->seed);
  return s->u[s->k < 1 ? s->k++ : 1];
}

static int has_truthiness(const orc_graph *g, uint64_t vid) {
  return !is_linear_zero(g->total_truthiness[vid]); /* variable.h:105 */
}

/* GibbsSamplerThread::sample_evid, src/gibbs_sampler.h:171-190 */
static uint64_t sample_evid(const orc_graph *g, const orc_opts *o, orc_state *s,
                            uint64_t vid, usrc *seed) {
  if (!o->noise_aware && g->is_evid[vid]) return g->assignment_dense[vid];
  if (o->noise_aware && has_truthiness(g, vid)) {
    double r = next_u(seed);
    double sum = 0;
    for (uint64_t i = 0; i < g->card[vid]; ++i) {
      sum += g->truthiness[g->var_val_base[vid] + i];
      if (sum >= r) return i;
    }
    return 0;
  }
  /* draw_sample evaluates the potentials first, then erand48: same stream */
  if (g->dtype[vid] == 0) {
    double pos = orc_potential(g, vid, 1, s->assign_evid, s->weights);
    double neg = orc_potential(g, vid, 0, s->assign_evid, s->weights);
    return orc_draw_boolean(pos, neg, next_u(seed));
  }
  return orc_draw_sample_r(g, vid, s->assign_evid, s->weights, next_u(seed));
}

/* src/gibbs_sampler.h:151-169 */
static void sample_single_variable(const orc_graph *g, const orc_opts *o, orc_state *s,
                                   uint64_t vid, usrc *seed) {
  if (g->is_evid[vid] && !o->sample_evidence) return;
  uint64_t p = orc_draw_sample_r(g, vid, s->assign_evid, s->weights, next_u(seed));
  if (p == ORC_INVALID) abort();
  s->assign_evid[vid] = p;
  ++s->nsamples[vid];
  if (g->dtype[vid] != 0 || p == 1)
    ++s->tallies[g->var_val_base[vid] + (g->dtype[vid] == 0 ? 0 : p)];
}

/* src/gibbs_sampler.h:127-149 */
static void sample_sgd_single_variable(const orc_graph *g, const orc_opts *o,
                                       orc_state *s, uint64_t vid, double stepsize,
                                       usrc *seed) {
  uint64_t p = orc_draw_sample_r(g, vid, s->assign_free, s->weights, next_u(seed));
  if (p == ORC_INVALID) abort();
  s->assign_free[vid] = p;
  uint64_t e = sample_evid(g, o, s, vid, seed);
  if (e == ORC_INVALID) abort();
  s->assign_evid[vid] = e;
  if (!o->learn_non_evidence && ((!o->noise_aware && !g->is_evid[vid]) ||
                                 (o->noise_aware && !has_truthiness(g, vid))))
    return;
  orc_sgd_on_variable(g, o, s, vid, stepsize);
}

void orc_sample_single_variable(const orc_graph *g, const orc_opts *o, orc_state *s,
                                uint64_t vid, unsigned short seed[3]) {
  usrc u = {seed, {0, 0}, 0};
  sample_single_variable(g, o, s, vid, &u);
}
void orc_sample_sgd_single_variable(const orc_graph *g, const orc_opts *o,
                                    orc_state *s, uint64_t vid, double stepsize,
                                    unsigned short seed[3]) {
  usrc u = {seed, {0, 0}, 0};
  sample_sgd_single_variable(g, o, s, vid, stepsize, &u);
}

/* The same per-variable steps in an explicit variable order with explicit uniforms
 * (r0[v]: first draw of v's step, r1[v]: second).  Sequential and in `order`, so it is the
 * reference algorithm for a different scan order -- what a chromatic / schedule-order
 * sweep must reproduce exactly. */
void orc_sweep_infer_seq(const orc_graph *g, const orc_opts *o, orc_state *s,
                         const uint64_t *order, uint64_t n, const double *r0) {
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t v = order[i];
    usrc u = {NULL, {r0[v], 0}, 0};
    sample_single_variable(g, o, s, v, &u);
  }
}
void orc_sweep_learn_seq(const orc_graph *g, const orc_opts *o, orc_state *s, double stepsize,
                         const uint64_t *order, uint64_t n, const double *r0, const double *r1) {
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t v = order[i];
    usrc u = {NULL, {r0[v], r1[v]}, 0};
    sample_sgd_single_variable(g, o, s, v, stepsize, &u);
  }
}

/* shard bounds, src/gibbs_sampler.cc:52-54 */
static void shard(uint64_t nvar, uint64_t i, uint64_t n, uint64_t *start, uint64_t *end) {
  *start = (nvar / n + 1) * i;
  *end = (nvar / n + 1) * (i + 1);
  if (*end > nvar) *end = nvar;
}

/* GibbsSamplerThread::sample, src/gibbs_sampler.cc:65-69 */
void orc_sweep_infer(const orc_graph *g, const orc_opts *o, orc_state *s,
                     uint64_t ith, uint64_t n, unsigned short seed[3]) {
  uint64_t a, b;
  shard(g->nvar, ith, n, &a, &b);
  usrc u = {seed, {0, 0}, 0};
  for (uint64_t v = a; v < b; ++v) sample_single_variable(g, o, s, v, &u);
}

/* GibbsSamplerThread::sample_sgd, src/gibbs_sampler.cc:71-75 */
void orc_sweep_learn(const orc_graph *g, const orc_opts *o, orc_state *s,
                     double stepsize, uint64_t ith, uint64_t n,
                     unsigned short seed[3]) {
  uint64_t a, b;
  shard(g->nvar, ith, n, &a, &b);
  usrc u = {seed, {0, 0}, 0};
  for (uint64_t v = a; v < b; ++v) sample_sgd_single_variable(g, o, s, v, stepsize, &u);
}

/* DimmWitted::learn + inference with one sampler, one worker:
 * src/dimmwitted.cc:162-207 (stepsize *= decay after each epoch :202),
 * :121-160 (clear tallies :127), seeds = rand() x3 (src/gibbs_sampler.cc:49). */
void orc_run(const orc_graph *g, const orc_opts *o, orc_state *s, uint64_t n_learn,
             uint64_t n_infer, double stepsize, double decay, long srand_seed) {
  if (srand_seed >= 0) srand((unsigned)srand_seed);
  unsigned short seed[3];
  /* set_random_seed(rand(), rand(), rand()): g++ on x86-64 evaluates the
   * arguments right to left, so the FIRST rand() lands in seed[2] [pinned by
   * tests/test_oracle_pin.py::test_whole_run_bit_for_bit] */
  seed[2] = (unsigned short)rand();
  seed[1] = (unsigned short)rand();
  seed[0] = (unsigned short)rand();
  double cur = stepsize;
  for (uint64_t e = 0; e < n_learn; ++e) {
    orc_sweep_learn(g, o, s, cur, 0, 1, seed);
    cur *= decay;
  }
  orc_clear_tallies(g, s);
  for (uint64_t e = 0; e < n_infer; ++e) orc_sweep_infer(g, o, s, 0, 1, seed);
}

/* ---- graph construction ------------------------------------------------- */

typedef struct {
  uint64_t value_dense, factor_id;
} tvf; /* TempValueFactor, src/variable.h:14-29 */

static int tvf_cmp(const void *pa, const void *pb) {
  const tvf *a = (const tvf *)pa, *b = (const tvf *)pb;
  if (a->value_dense != b->value_dense) return a->value_dense < b->value_dense ? -1 : 1;
  if (a->factor_id != b->factor_id) return a->factor_id < b->factor_id ? -1 : 1;
  return 0;
}

#define NEW(T, n) ((T *)calloc((n) ? (n) : 1, sizeof(T)))

void orc_graph_free(orc_graph *g) {
  if (!g) return;
  free(g->dtype); free(g->is_evid); free(g->card); free(g->assignment_dense);
  free(g->var_val_base); free(g->total_truthiness); free(g->value);
  free(g->truthiness); free(g->fi_base); free(g->fi_len); free(g->factor_index);
  free(g->func); free(g->num_vars); free(g->vif_base); free(g->wid); free(g->fv);
  free(g->vif_vid); free(g->vif_eq); free(g->w_init); free(g->w_fixed);
  free(g);
}

orc_graph *orc_graph_build(uint64_t nvar, const uint64_t *var_rec, uint64_t ndom,
                           const uint64_t *dom_vid, const uint64_t *dom_ptr,
                           const uint64_t *dom_val, const double *dom_truth,
                           uint64_t nfac, const int32_t *func, const uint64_t *fac_ptr,
                           const uint64_t *edge_vid, const uint64_t *edge_eq,
                           const uint64_t *fac_wid, const double *fac_fv,
                           uint64_t nweight, const double *w_init,
                           const uint8_t *w_fixed) {
  orc_graph *g = NEW(orc_graph, 1);
  uint64_t nedge = fac_ptr[nfac];
  g->nvar = nvar; g->nfac = nfac; g->nweight = nweight; g->nedge = nedge;
  g->dtype = NEW(uint8_t, nvar); g->is_evid = NEW(uint8_t, nvar);
  g->card = NEW(uint64_t, nvar); g->assignment_dense = NEW(uint64_t, nvar);
  g->var_val_base = NEW(uint64_t, nvar); g->total_truthiness = NEW(double, nvar);
  g->func = NEW(int32_t, nfac); g->num_vars = NEW(uint64_t, nfac);
  g->vif_base = NEW(uint64_t, nfac); g->wid = NEW(uint64_t, nfac);
  g->fv = NEW(double, nfac);
  g->vif_vid = NEW(uint64_t, nedge); g->vif_eq = NEW(uint64_t, nedge);
  g->w_init = NEW(double, nweight); g->w_fixed = NEW(uint8_t, nweight);
  int64_t *dom_of = NEW(int64_t, nvar); /* domain block of each var, -1 none */
  uint64_t *deg = NEW(uint64_t, nvar + 1);
  tvf *adj = NULL;
  int ok = 1;

  memcpy(g->w_init, w_init, sizeof(double) * nweight);
  memcpy(g->w_fixed, w_fixed, nweight);

  /* load_variables, src/binary_format.cc:72-126 */
  for (uint64_t v = 0; v < nvar; ++v) {
    uint64_t role = var_rec[4 * v], init = var_rec[4 * v + 1];
    uint64_t dt = var_rec[4 * v + 2];
    if (dt > 1) ok = 0; /* :101-105 aborts */
    g->dtype[v] = (uint8_t)dt;
    g->is_evid[v] = role >= 1;                 /* :107 */
    g->assignment_dense[v] = role >= 1 ? init : 0; /* :109 */
    g->card[v] = var_rec[4 * v + 3];
    dom_of[v] = -1;
  }
  /* load_domains, src/binary_format.cc:192-226: index = position in block */
  for (uint64_t d = 0; ok && d < ndom; ++d) {
    uint64_t v = dom_vid[d];
    if (v >= nvar || g->dtype[v] == 0 || g->card[v] != dom_ptr[d + 1] - dom_ptr[d]) {
      ok = 0; /* asserts :206-207 */
      break;
    }
    dom_of[v] = (int64_t)d;
    if (g->assignment_dense[v]) { /* :220 quirk: value 0 is never remapped */
      uint64_t found = ORC_INVALID;
      /* unordered_map insert-or-assign :215: a repeated value keeps its LAST index */
      for (uint64_t i = dom_ptr[d]; i < dom_ptr[d + 1]; ++i)
        if (dom_val[i] == g->assignment_dense[v]) found = i - dom_ptr[d];
      if (found == ORC_INVALID) ok = 0; /* .at() throws */
      g->assignment_dense[v] = found;
    }
  }
  /* load_factors, src/binary_format.cc:128-190 (single stream => ids in order) */
  for (uint64_t f = 0; ok && f < nfac; ++f) {
    g->func[f] = func[f];
    g->num_vars[f] = fac_ptr[f + 1] - fac_ptr[f];
    g->vif_base[f] = fac_ptr[f];
    g->wid[f] = fac_wid[f];
    g->fv[f] = fac_fv[f];
    for (uint64_t e = fac_ptr[f]; e < fac_ptr[f + 1]; ++e) {
      uint64_t v = edge_vid[e];
      if (v >= nvar) { ok = 0; break; } /* assert :159 */
      uint64_t dense = edge_eq[e];      /* get_domain_index, variable.h:118 */
      if (dom_of[v] >= 0) {
        uint64_t d = (uint64_t)dom_of[v], found = ORC_INVALID;
        for (uint64_t i = dom_ptr[d]; i < dom_ptr[d + 1]; ++i)
          if (dom_val[i] == edge_eq[e]) found = i - dom_ptr[d];
        if (found == ORC_INVALID) { ok = 0; break; }
        dense = found;
      }
      g->vif_vid[e] = v;
      g->vif_eq[e] = dense;
      ++deg[v + 1];
    }
  }
  if (!ok) goto fail;

  /* adjacency lists (Variable::add_value_factor, variable.h:93-98) */
  for (uint64_t v = 0; v < nvar; ++v) deg[v + 1] += deg[v];
  adj = NEW(tvf, nedge);
  {
    uint64_t *fill = NEW(uint64_t, nvar);
    for (uint64_t f = 0; f < nfac; ++f)
      for (uint64_t e = fac_ptr[f]; e < fac_ptr[f + 1]; ++e) {
        uint64_t v = g->vif_vid[e];
        tvf *t = &adj[deg[v] + fill[v]++];
        t->value_dense = g->dtype[v] == 0 ? 0 : g->vif_eq[e]; /* :166-169 */
        t->factor_id = f;
      }
    free(fill);
  }

  /* construct_index(_part), src/factor_graph.cc:90-199 */
  uint64_t nval = 0;
  for (uint64_t v = 0; v < nvar; ++v) nval += g->dtype[v] == 0 ? 1 : g->card[v];
  g->nval = nval;
  g->value = NEW(uint64_t, nval); g->truthiness = NEW(double, nval);
  g->fi_base = NEW(uint64_t, nval); g->fi_len = NEW(uint64_t, nval);
  g->factor_index = NEW(uint64_t, nedge);
  uint64_t vb = 0, fb = 0;
  for (uint64_t v = 0; v < nvar; ++v) {
    g->var_val_base[v] = vb;
    g->total_truthiness[v] = 0;
    if (g->dtype[v] == 0) {
      g->value[vb++] = 0; /* :146-149 */
    } else if (dom_of[v] >= 0) {
      uint64_t d = (uint64_t)dom_of[v]; /* :151-164 */
      for (uint64_t j = 0; j < g->card[v]; ++j) {
        g->value[vb] = dom_val[dom_ptr[d] + j];
        g->truthiness[vb] = dom_truth[dom_ptr[d] + j];
        g->total_truthiness[v] += dom_truth[dom_ptr[d] + j];
        ++vb;
      }
    } else {
      for (uint64_t j = 0; j < g->card[v]; ++j) g->value[vb++] = j; /* :169-172 */
    }
    uint64_t n = deg[v + 1] - deg[v];
    if (n) { /* :177-197 */
      tvf *a = adj + deg[v];
      qsort(a, n, sizeof(tvf), tvf_cmp);
      uint64_t value_dense = ORC_INVALID, last = ORC_INVALID;
      for (uint64_t i = 0; i < n; ++i) {
        if (a[i].value_dense != value_dense) {
          value_dense = a[i].value_dense;
          uint64_t icard = g->dtype[v] == 0 ? g->card[v] : g->card[v];
          if (value_dense >= icard) { ok = 0; break; } /* assert :186 */
          g->fi_base[g->var_val_base[v] + value_dense] = fb;
        } else if (a[i].factor_id == last) {
          continue;
        }
        g->factor_index[fb++] = a[i].factor_id;
        ++g->fi_len[g->var_val_base[v] + value_dense];
        last = a[i].factor_id;
      }
      if (!ok) break;
    }
  }
  g->nindex = fb;
  if (!ok) goto fail;
  free(adj); free(deg); free(dom_of);
  return g;
fail:
  free(adj); free(deg); free(dom_of);
  orc_graph_free(g);
  return NULL;
}

/* InferenceResult ctor, src/inference_result.cc:24-42 */
orc_state *orc_state_new(const orc_graph *g) {
  orc_state *s = NEW(orc_state, 1);
  s->assign_free = NEW(uint64_t, g->nvar);
  s->assign_evid = NEW(uint64_t, g->nvar);
  s->weights = NEW(double, g->nweight);
  s->tallies = NEW(uint64_t, g->nval);
  s->nsamples = NEW(uint64_t, g->nvar);
  memcpy(s->weights, g->w_init, sizeof(double) * g->nweight);
  for (uint64_t v = 0; v < g->nvar; ++v)
    s->assign_free[v] = s->assign_evid[v] = g->is_evid[v] ? g->assignment_dense[v] : 0;
  return s;
}

void orc_state_free(orc_state *s) {
  if (!s) return;
  free(s->assign_free); free(s->assign_evid); free(s->weights);
  free(s->tallies); free(s->nsamples); free(s);
}

/* src/inference_result.cc:107-114 */
void orc_clear_tallies(const orc_graph *g, orc_state *s) {
  memset(s->tallies, 0, sizeof(uint64_t) * g->nval);
  memset(s->nsamples, 0, sizeof(uint64_t) * g->nvar);
}
