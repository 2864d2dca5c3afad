// Test-infrastructure (NOT product code): extern "C" window onto the UNMODIFIED
// reference objects (compiled from /root/reference/src by oracle/Makefile at -O2,
// IEEE-exact) so that tests can pin the C oracle (oracle/dw_oracle.c) and the CUDA
// path function-by-function against the reference itself.  Nothing in
// sampler_b200/ links or loads this.
//
// Every entry point names the reference function it forwards to.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "dimmwitted.h"
#include "factor.h"
#include "factor_graph.h"
#include "gibbs_sampler.h"
#include "binary_format.h"

using namespace dd;

namespace {
struct RefHandle {
  std::unique_ptr<CmdParser> opts;
  std::unique_ptr<FactorGraph> fg;
  std::unique_ptr<InferenceResult> infrs;
  std::unique_ptr<GibbsSamplerThread> sampler;
};
}  // namespace

extern "C" {

// --- stand-alone math -------------------------------------------------------

// src/common.h:118-132
double ref_logadd(double a, double b) { return logadd(a, b); }

// Factor::potential, src/factor.h:59-86, on an ad-hoc factor (vids/eqs given).
double ref_factor_potential(int func_id, uint64_t arity, const uint64_t *vids,
                            const uint64_t *eqs, double feature_value,
                            const uint64_t *assignments, uint64_t vid,
                            uint64_t proposal) {
  std::vector<FactorToVariable> vifs(arity);
  for (uint64_t i = 0; i < arity; ++i) vifs[i] = FactorToVariable(vids[i], eqs[i]);
  Factor f(0, feature_value, 0, (FACTOR_FUNCTION_TYPE)func_id, arity);
  f.vif_base = 0;
  return f.potential(vifs.data(), (const size_t *)assignments, vid, proposal);
}

// glibc erand48 stream the reference's draws consume (src/gibbs_sampler.h:203,237)
void ref_uniforms(unsigned short s0, unsigned short s1, unsigned short s2, int n,
                  double *out) {
  unsigned short s[3] = {s0, s1, s2};
  for (int i = 0; i < n; ++i) out[i] = erand48(s);
}

// --- whole-graph handle: mirrors dd::gibbs() src/dimmwitted.cc:55-72 and the
// test fixture test/sampler_test.cc:23-49 -------------------------------------

void *ref_open(int argc, const char *const argv[]) {
  RefHandle *h = new RefHandle;
  h->opts.reset(new CmdParser(argc, argv));
  FactorGraphDescriptor meta = read_meta(h->opts->fg_file);
  h->fg.reset(new FactorGraph(meta));
  h->fg->load_variables(h->opts->variable_file);
  h->fg->load_weights(h->opts->weight_file);
  h->fg->load_domains(h->opts->domain_file);
  h->fg->load_factors(h->opts->factor_file);
  h->fg->safety_check();
  h->fg->construct_index();
  h->infrs.reset(new InferenceResult(*h->fg, h->fg->weights.get(), *h->opts));
  h->sampler.reset(new GibbsSamplerThread(*h->fg, *h->infrs, 0, 1, *h->opts));
  return h;
}
void ref_close(void *p) { delete (RefHandle *)p; }

#define H ((RefHandle *)p)
uint64_t ref_num_variables(void *p) { return H->fg->size.num_variables; }
uint64_t ref_num_factors(void *p) { return H->fg->size.num_factors; }
uint64_t ref_num_weights(void *p) { return H->fg->size.num_weights; }
uint64_t ref_num_edges(void *p) { return H->fg->size.num_edges; }
uint64_t ref_num_values(void *p) { return H->fg->size.num_values; }

// graph structure after construct_index (src/factor_graph.cc:90-199)
// vars: [nvar][6] = dtype, is_evid, cardinality, assignment_dense, var_val_base, has_truthiness
void ref_get_variables(void *p, uint64_t *out) {
  for (size_t i = 0; i < H->fg->size.num_variables; ++i) {
    const Variable &v = H->fg->variables[i];
    out[6 * i + 0] = v.domain_type;
    out[6 * i + 1] = v.is_evid;
    out[6 * i + 2] = v.cardinality;
    out[6 * i + 3] = v.assignment_dense;
    out[6 * i + 4] = v.var_val_base;
    out[6 * i + 5] = v.has_truthiness();
  }
}
// values: value[nval], truthiness[nval], base[nval], len[nval]
void ref_get_values(void *p, uint64_t *value, double *truthiness, uint64_t *base,
                    uint64_t *len) {
  for (size_t i = 0; i < H->fg->size.num_values; ++i) {
    const VariableToFactor &v = H->fg->values[i];
    value[i] = v.value;
    truthiness[i] = v.truthiness;
    base[i] = v.factor_index_base;
    len[i] = v.factor_index_length;
  }
}
void ref_get_factor_index(void *p, uint64_t *out, uint64_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = H->fg->factor_index[i];
}
// factors: [nfac][4] = func_id, num_vars, vif_base, weight_id ; fv[nfac]
void ref_get_factors(void *p, int64_t *out, double *fv) {
  for (size_t i = 0; i < H->fg->size.num_factors; ++i) {
    const Factor &f = H->fg->factors[i];
    out[4 * i + 0] = f.func_id;
    out[4 * i + 1] = f.num_vars;
    out[4 * i + 2] = f.vif_base;
    out[4 * i + 3] = f.weight_id;
    fv[i] = f.feature_value;
  }
}
void ref_get_vifs(void *p, uint64_t *vid, uint64_t *eq) {
  for (size_t i = 0; i < H->fg->size.num_edges; ++i) {
    vid[i] = H->fg->vifs[i].vid;
    eq[i] = H->fg->vifs[i].dense_equal_to;
  }
}

// state (src/inference_result.h:27-43)
static size_t *chain(RefHandle *h, int which) {
  return which ? h->infrs->assignments_evid.get() : h->infrs->assignments_free.get();
}
void ref_get_assignments(void *p, int which, uint64_t *out) {
  std::memcpy(out, chain(H, which), sizeof(uint64_t) * H->infrs->nvars);
}
void ref_set_assignments(void *p, int which, const uint64_t *in) {
  std::memcpy(chain(H, which), in, sizeof(uint64_t) * H->infrs->nvars);
}
void ref_get_weights(void *p, double *out) {
  std::memcpy(out, H->infrs->weight_values.get(), sizeof(double) * H->infrs->nweights);
}
void ref_set_weights(void *p, const double *in) {
  std::memcpy(H->infrs->weight_values.get(), in, sizeof(double) * H->infrs->nweights);
}
void ref_get_tallies(void *p, uint64_t *tallies, uint64_t *nsamples) {
  std::memcpy(tallies, H->infrs->sample_tallies.get(), sizeof(uint64_t) * H->infrs->ntallies);
  std::memcpy(nsamples, H->infrs->agg_nsamples.get(), sizeof(uint64_t) * H->infrs->nvars);
}
void ref_clear_tallies(void *p) { H->infrs->clear_variabletally(); }

void ref_set_seed(void *p, unsigned short s0, unsigned short s1, unsigned short s2) {
  H->sampler->set_random_seed(s0, s1, s2);  // src/gibbs_sampler.cc:57-63
}

// FactorGraph::potential, src/factor_graph.h:127-145
double ref_potential(void *p, uint64_t vid, uint64_t proposal, int which) {
  return H->fg->potential(H->fg->variables[vid], proposal, chain(H, which),
                          H->infrs->weight_values.get());
}
// Factor::potential of an indexed factor, src/factor.h:59
double ref_factor_potential_by_id(void *p, uint64_t fid, int which, uint64_t vid,
                                  uint64_t proposal) {
  return H->fg->factors[fid].potential(H->fg->vifs.get(), chain(H, which), vid, proposal);
}
// GibbsSamplerThread::draw_sample, src/gibbs_sampler.h:192-254
uint64_t ref_draw_sample(void *p, uint64_t vid, int which) {
  return H->sampler->draw_sample(H->fg->variables[vid], chain(H, which),
                                 H->infrs->weight_values.get());
}
// src/gibbs_sampler.h:171-190
uint64_t ref_sample_evid(void *p, uint64_t vid) {
  return H->sampler->sample_evid(H->fg->variables[vid]);
}
// src/gibbs_sampler.h:151-169
void ref_sample_single_variable(void *p, uint64_t vid) {
  H->sampler->sample_single_variable(vid);
}
// src/gibbs_sampler.h:127-149
void ref_sample_sgd_single_variable(void *p, uint64_t vid, double stepsize) {
  H->sampler->sample_sgd_single_variable(vid, stepsize);
}
// src/factor_graph.cc:262-314
void ref_sgd_on_variable(void *p, uint64_t vid, double stepsize, int noise_aware) {
  H->fg->sgd_on_variable(H->fg->variables[vid], *H->infrs, stepsize, noise_aware);
}
// src/inference_result.h:66-85
void ref_update_weight(void *p, uint64_t wid, double stepsize, double gradient) {
  H->infrs->update_weight(wid, stepsize, gradient);
}
// src/gibbs_sampler.cc:65-75 with one shard (-t 1)
void ref_sweep_infer(void *p) { H->sampler->sample(); }
void ref_sweep_learn(void *p, double stepsize) { H->sampler->sample_sgd(stepsize); }
#undef H

}  // extern "C"
