/* dw_oracle.h -- CPU ORACLE for the DimmWitted Gibbs hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * nothing under sampler_b200/ links, loads or calls it, and the product fails
 * loudly when its CUDA library is missing rather than falling back to this.
 *
 * A plain-C, single-threaded restatement of the reference algorithm
 * (HazyResearch/sampler @ 3651366), each function citing the reference
 * file:line it follows.  Parity is PINNED: tests/test_oracle_*.py check it
 * against the reference's own known-answer tests (test/factor_test.cc,
 * test/factor_graph_test.cc, test/sampler_test.cc, committed as
 * tests/golden/) and, where oracle/_ref/libdwref.so was built from the
 * reference sources, against the reference itself function-by-function and
 * bit-for-bit on whole `-t 1 -c 1` runs.
 */
#ifndef DW_ORACLE_H_
#define DW_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/common.h:33-46 */
enum {
  ORC_FUNC_IMPLY_NATURAL = 0,
  ORC_FUNC_OR = 1,
  ORC_FUNC_AND = 2,
  ORC_FUNC_EQUAL = 3,
  ORC_FUNC_ISTRUE = 4,
  ORC_FUNC_LINEAR = 7,
  ORC_FUNC_RATIO = 8,
  ORC_FUNC_LOGICAL = 9,
  ORC_FUNC_AND_CATEGORICAL = 12,
  ORC_FUNC_IMPLY_MLN = 13
};

#define ORC_INVALID ((uint64_t)-1)

/* Indexed factor graph = the reference's FactorGraph after construct_index()
 * (src/factor_graph.h:58-70), as parallel arrays. */
typedef struct orc_graph {
  uint64_t nvar, nfac, nweight, nedge, nval, nindex;
  /* Variable, src/variable.h:38-121 */
  uint8_t *dtype;    /* 0 Boolean, 1 categorical */
  uint8_t *is_evid;
  uint64_t *card;
  uint64_t *assignment_dense;
  uint64_t *var_val_base;
  double *total_truthiness;
  /* VariableToFactor, src/variable.h:123-142 */
  uint64_t *value;
  double *truthiness;
  uint64_t *fi_base, *fi_len;
  uint64_t *factor_index;
  /* Factor, src/factor.h:21-26 */
  int32_t *func;
  uint64_t *num_vars, *vif_base, *wid;
  double *fv;
  /* FactorToVariable, src/variable.h:147-167 */
  uint64_t *vif_vid, *vif_eq;
  /* Weight, src/weight.h:13-25 */
  double *w_init;
  uint8_t *w_fixed;
} orc_graph;

/* InferenceResult, src/inference_result.h:27-43 */
typedef struct orc_state {
  uint64_t *assign_free, *assign_evid;
  double *weights;
  uint64_t *tallies, *nsamples;
} orc_state;

/* the CmdParser fields the hot path reads, src/cmd_parser.h:41-56 */
typedef struct orc_opts {
  int sample_evidence, learn_non_evidence, noise_aware, reg_l1;
  double reg_param;
} orc_opts;

/* Build from decoded records; mirrors load_variables/load_domains/load_factors
 * (src/binary_format.cc:72-226) + construct_index (src/factor_graph.cc:90-199).
 * var_rec: [nvar][4] = role, initial_value, dtype, cardinality (record order = id).
 * domains: ndom blocks; dom_vid[i], values dom_val[dom_ptr[i]..dom_ptr[i+1]).
 * factors: fac_ptr CSR over edges; edge_eq holds the ORIGINAL (sparse) equal_to.
 * Returns NULL on malformed input (what the reference would assert on). */
orc_graph *orc_graph_build(uint64_t nvar, const uint64_t *var_rec, uint64_t ndom,
                           const uint64_t *dom_vid, const uint64_t *dom_ptr,
                           const uint64_t *dom_val, const double *dom_truth,
                           uint64_t nfac, const int32_t *func, const uint64_t *fac_ptr,
                           const uint64_t *edge_vid, const uint64_t *edge_eq,
                           const uint64_t *fac_wid, const double *fac_fv,
                           uint64_t nweight, const double *w_init,
                           const uint8_t *w_fixed);
void orc_graph_free(orc_graph *g);

orc_state *orc_state_new(const orc_graph *g);
void orc_state_free(orc_state *s);
void orc_clear_tallies(const orc_graph *g, orc_state *s);

/* math */
double orc_logadd(double a, double b);
double orc_factor_potential(int func, uint64_t n, const uint64_t *vids,
                            const uint64_t *eqs, double fv,
                            const uint64_t *assignments, uint64_t vid,
                            uint64_t proposal);
double orc_potential(const orc_graph *g, uint64_t vid, uint64_t proposal,
                     const uint64_t *assignments, const double *weights);
uint64_t orc_draw_boolean(double pot_pos, double pot_neg, double r);
uint64_t orc_draw_categorical(const double *pots, uint64_t card, double r);
/* draw with an explicit uniform */
uint64_t orc_draw_sample_r(const orc_graph *g, uint64_t vid,
                           const uint64_t *assignments, const double *weights,
                           double r);
void orc_update_weight(const orc_opts *o, double *weights, uint64_t wid,
                       double stepsize, double gradient);
void orc_sgd_on_variable(const orc_graph *g, const orc_opts *o, orc_state *s,
                         uint64_t vid, double stepsize);

/* per-variable steps drawing from an erand48 state (glibc, like the reference) */
void orc_sample_single_variable(const orc_graph *g, const orc_opts *o, orc_state *s,
                                uint64_t vid, unsigned short seed[3]);
void orc_sample_sgd_single_variable(const orc_graph *g, const orc_opts *o,
                                    orc_state *s, uint64_t vid, double stepsize,
                                    unsigned short seed[3]);
/* one shard of a sweep, src/gibbs_sampler.cc:52-54,65-75 */
void orc_sweep_infer(const orc_graph *g, const orc_opts *o, orc_state *s,
                     uint64_t ith_shard, uint64_t n_shards, unsigned short seed[3]);
void orc_sweep_learn(const orc_graph *g, const orc_opts *o, orc_state *s,
                     double stepsize, uint64_t ith_shard, uint64_t n_shards,
                     unsigned short seed[3]);
/* explicit scan order + explicit uniforms (r0[v] first draw, r1[v] second draw of v's step) */
void orc_sweep_infer_seq(const orc_graph *g, const orc_opts *o, orc_state *s,
                         const uint64_t *order, uint64_t n, const double *r0);
void orc_sweep_learn_seq(const orc_graph *g, const orc_opts *o, orc_state *s, double stepsize,
                         const uint64_t *order, uint64_t n, const double *r0, const double *r1);
/* whole single-replica, single-thread schedule = `dw gibbs -t 1 -c 1` after
 * srand(srand_seed) (src/dimmwitted.cc:121-207); srand_seed<0 => no srand. */
void orc_run(const orc_graph *g, const orc_opts *o, orc_state *s, uint64_t n_learn,
             uint64_t n_infer, double stepsize, double decay, long srand_seed);

#ifdef __cplusplus
}
#endif
#endif
