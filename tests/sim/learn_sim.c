/* learn_sim.c -- TEST INFRASTRUCTURE: a CPU model of the GPU learning schedule on top of the
 * oracle's per-variable functions (oracle/dw_oracle.c), used to choose / check the design of the
 * parallel learning kernels against the reference's -t 1 goldens without a GPU.
 *
 * Model ("local replicas"): R replicas (= CTAs), each with a private live copy of the weights;
 * every replica has WPC warps; the 32-variable slices of a colour class are dealt round-robin to
 * the R*WPC warps; a warp draws its 32 variables with the replica's weights as they are at that
 * moment (phase A) and then applies the 32 SGD steps one after the other in lane order to the
 * replica's copy (phase B, update_weight verbatim).  At the end of a class (merge_per_class) or
 * of the epoch the copies are merged:  w = w0 + sum_r n_r/(sum n) (w_r - w0)  per weight, n_r =
 * updates replica r made to that weight. */
#include "../../oracle/dw_oracle.c"

typedef struct {
  int R, WPC, merge_per_class, batch_cta; /* batch_cta: phase A of all warps of a CTA before any phase B */
  int merge_mode; /* 0 weighted average, 1 plain average, 2 sum of deltas */
  int SW; /* slice width (32 on the GPU) */
} sim_cfg;

static void slice_phase_a(const orc_graph *g, const orc_opts *o, orc_state *s, const uint64_t *vars, int n,
                          unsigned short seed[3]) {
  for (int i = 0; i < n; ++i) {
    uint64_t vid = vars[i];
    usrc u = {seed, {0, 0}, 0};
    uint64_t p = orc_draw_sample_r(g, vid, s->assign_free, s->weights, next_u(&u));
    s->assign_free[vid] = p;
    s->assign_evid[vid] = sample_evid(g, o, s, vid, &u);
  }
}
static const double *g_step_scale = 0; /* per-weight step scale, or NULL */
void sim_set_step_scale(const double *p) { g_step_scale = p; }
static void slice_phase_b(const orc_graph *g, const orc_opts *o, orc_state *s, const uint64_t *vars, int n,
                          double stepsize) {
  for (int i = 0; i < n; ++i) {
    uint64_t vid = vars[i];
    if (!o->learn_non_evidence &&
        ((!o->noise_aware && !g->is_evid[vid]) || (o->noise_aware && !has_truthiness(g, vid))))
      continue;
    if (!g_step_scale || g->dtype[vid] != 0) { orc_sgd_on_variable(g, o, s, vid, stepsize); continue; }
    uint64_t row = g->var_val_base[vid];
    for (uint64_t j = 0; j < g->fi_len[row]; ++j) {
      uint64_t fid = g->factor_index[g->fi_base[row] + j];
      if (g->w_fixed[g->wid[fid]]) continue;
      double pe = factor_pot(g, fid, s->assign_evid, vid, g->assignment_dense[vid]);
      double pf = factor_pot(g, fid, s->assign_free, ORC_INVALID, ORC_INVALID);
      orc_update_weight(o, s->weights, g->wid[fid], stepsize * g_step_scale[g->wid[fid]], pf - pe);
    }
  }
}

/* counts updates per weight: wrap by diffing is not possible, so count structurally */
static void count_updates(const orc_graph *g, const orc_opts *o, const orc_state *s, const uint64_t *vars, int n,
                          double *cnt) {
  for (int i = 0; i < n; ++i) {
    uint64_t vid = vars[i];
    if (!o->learn_non_evidence &&
        ((!o->noise_aware && !g->is_evid[vid]) || (o->noise_aware && !has_truthiness(g, vid))))
      continue;
    if (g->dtype[vid] == 0) {
      uint64_t row = g->var_val_base[vid];
      for (uint64_t j = 0; j < g->fi_len[row]; ++j) {
        uint64_t f = g->factor_index[g->fi_base[row] + j];
        if (!g->w_fixed[g->wid[f]]) cnt[g->wid[f]] += 1;
      }
    } else {
      uint64_t row = g->var_val_base[vid] + g->assignment_dense[vid];
      for (uint64_t j = 0; j < g->fi_len[row]; ++j) {
        uint64_t f = g->factor_index[g->fi_base[row] + j];
        if (!g->w_fixed[g->wid[f]]) cnt[g->wid[f]] += 1;
      }
    }
  }
}

static void merge(const orc_graph *g, const sim_cfg *c, double *W, double *cnt, double *w0, double *wglob) {
  uint64_t nw = g->nweight;
  for (uint64_t k = 0; k < nw; ++k) {
    double tot = 0, acc = 0;
    for (int r = 0; r < c->R; ++r) tot += cnt[(uint64_t)r * nw + k];
    if (c->merge_mode == 1) {
      for (int r = 0; r < c->R; ++r) acc += W[(uint64_t)r * nw + k];
      wglob[k] = acc / c->R;
    } else if (c->merge_mode == 2) {
      for (int r = 0; r < c->R; ++r) acc += W[(uint64_t)r * nw + k] - w0[k];
      wglob[k] = w0[k] + acc;
    } else if (tot > 0) {
      for (int r = 0; r < c->R; ++r) acc += cnt[(uint64_t)r * nw + k] / tot * (W[(uint64_t)r * nw + k] - w0[k]);
      wglob[k] = w0[k] + acc;
    } else wglob[k] = w0[k];
    for (int r = 0; r < c->R; ++r) { W[(uint64_t)r * nw + k] = wglob[k]; cnt[(uint64_t)r * nw + k] = 0; }
    w0[k] = wglob[k];
  }
}

/* one learning launch over explicit slices: slice i = order[sl_ptr[i] .. sl_ptr[i+1]), visited in
 * index order by G = R*WPC warps round-robin; step_scale[k] scales the step of weight k (NULL = 1) */
void sim_learn_launch(const orc_graph *g, const orc_opts *o, orc_state *s, double stepsize, const uint64_t *order,
                      const uint64_t *sl_ptr, uint64_t nsl, const sim_cfg *c, double *W, double *cnt, double *w0,
                      unsigned short seed[3], int do_merge) {
  uint64_t nw = g->nweight;
  double *glob = s->weights;
  int G = c->R * c->WPC;
  for (uint64_t t = 0; t * G < nsl; ++t) {
    for (int r = 0; r < c->R; ++r) {
      s->weights = W + (uint64_t)r * nw;
      for (int pass = 0; pass < (c->batch_cta ? 2 : 1); ++pass)
        for (int w = 0; w < c->WPC; ++w) {
          uint64_t sl = t * G + (uint64_t)r * c->WPC + w;
          if (sl >= nsl) continue;
          const uint64_t *vars = order + sl_ptr[sl];
          int n = (int)(sl_ptr[sl + 1] - sl_ptr[sl]);
          if (!c->batch_cta || pass == 0) slice_phase_a(g, o, s, vars, n, seed);
          if (!c->batch_cta || pass == 1) {
            count_updates(g, o, s, vars, n, cnt + (uint64_t)r * nw);
            slice_phase_b(g, o, s, vars, n, stepsize);
          }
        }
    }
  }
  s->weights = glob;
  if (do_merge) merge(g, c, W, cnt, w0, glob);
}

/* plain inference sweeps in the given order with erand48 */
void sim_infer(const orc_graph *g, const orc_opts *o, orc_state *s, const uint64_t *order, uint64_t n, int sweeps,
               unsigned short seed[3]) {
  for (int k = 0; k < sweeps; ++k)
    for (uint64_t i = 0; i < n; ++i) {
      usrc u = {seed, {0, 0}, 0};
      sample_single_variable(g, o, s, order[i], &u);
    }
}
