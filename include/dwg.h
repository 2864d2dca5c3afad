/* dwg.h -- C ABI of the B200-native DimmWitted Gibbs hot path (libdwg.so).
 *
 * The reference (HazyResearch/sampler @ 3651366) has no plugin/FFI interface; its
 * seam for this path is the DimmWitted / GibbsSampler / InferenceResult class
 * boundary.  Each entry point below names the reference interface it replaces
 * (file:line relative to the reference tree).  Plain pointers and sizes only, no
 * C++/torch types; every call returns 0 or a negative DWG_ERR_* and never throws
 * or aborts (the reference asserts/aborts: src/factor.h:84, src/binary_format.h:110);
 * dwg_last_error() gives the message for the calling thread.
 *
 * Objects
 *   dwg_graph  : a factor graph compiled on the host into the device layout
 *                (== FactorGraph after load_* + safety_check + construct_index,
 *                src/dimmwitted.cc:63-72).  Immutable, shareable between replicas.
 *   dwg_handle : one replica resident on one GPU: graph copy + both Gibbs chains +
 *                weights + tallies (== one GibbsSampler with its FactorGraph copy and
 *                InferenceResult, src/gibbs_sampler.h:18-57, src/dimmwitted.cc:97-119).
 * Ids crossing the ABI are always the ORIGINAL ids/values of the input files; the
 * permutation into schedule order is internal.
 */
#ifndef DWG_H_
#define DWG_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DWG_OK 0
#define DWG_ERR_INVALID (-1)     /* malformed input: what the reference asserts on */
#define DWG_ERR_IO (-2)          /* cannot open/read a file */
#define DWG_ERR_CUDA (-3)        /* CUDA runtime error (no device, OOM, launch failure) */
#define DWG_ERR_UNSUPPORTED (-4) /* beyond the 32-bit device layout, see DESIGN.md */
#define DWG_ERR_NCCL (-5)

/* FACTOR_FUNCTION_TYPE, src/common.h:33-46 */
enum {
  DWG_FUNC_IMPLY_NATURAL = 0, DWG_FUNC_OR = 1, DWG_FUNC_AND = 2, DWG_FUNC_EQUAL = 3,
  DWG_FUNC_ISTRUE = 4, DWG_FUNC_LINEAR = 7, DWG_FUNC_RATIO = 8, DWG_FUNC_LOGICAL = 9,
  DWG_FUNC_AND_CATEGORICAL = 12, DWG_FUNC_IMPLY_MLN = 13
};

/* sweep schedules */
enum {
  DWG_SCHED_AUTO = 0,      /* sequential for tiny graphs (<= DWG_SEQUENTIAL_MAX_EDGES edge records),
                              else hogwild if every factor touches one variable, else chromatic */
  DWG_SCHED_HOGWILD = 1,   /* one launch per sweep, all variables concurrently (like the
                              reference's Hogwild threads, src/gibbs_sampler.cc:20-33) */
  DWG_SCHED_CHROMATIC = 2, /* host greedy colouring; one launch per colour class: exact Gibbs */
  DWG_SCHED_SEQUENTIAL = 3 /* one chain walking the variables in id order with the reference's
                              non-atomic per-sample weight update: `dw -t 1 -c 1` semantics exactly.
                              For graphs too small to parallelise, where per-sample SGD with large
                              steps and a per-launch mini-batch visibly differ. */
};
#define DWG_SEQUENTIAL_MAX_EDGES 256

/* The CmdParser fields the hot path reads (src/cmd_parser.h:41-56) + device knobs. */
typedef struct dwg_opts {
  int sample_evidence;     /* --sample_evidence    */
  int learn_non_evidence;  /* --learn_non_evidence */
  int noise_aware;         /* --noise_aware        */
  int regularization;      /* 0 = l2 (default), 1 = l1; --regularization */
  double reg_param;        /* -b / --reg_param     */
  int schedule;            /* DWG_SCHED_*          */
  uint64_t seed;           /* counter-based RNG key (the reference: rand(), gibbs_sampler.cc:49) */
  uint32_t replica;        /* RNG stream of this replica (== nodeid, gibbs_sampler.h:34) */
  int learn_chunks;        /* categorical graphs (CSR learning kernels) only: batch weight updates per colour
                              class and learning sweep; 0 = auto (about 32 updates per sweep).  Boolean
                              graphs learn with live weights and ignore it */
  int defer_upload;        /* dwg_create allocates the replica's device arrays but does not fill the graph
                              segments: another replica's copy arrives over NVLink (dwg_graph_segments +
                              a broadcast), then dwg_reset_state.  The reference deep-copies the graph per
                              NUMA node from replica 0 (src/dimmwitted.cc:113-116) */
} dwg_opts_t;
void dwg_opts_default(dwg_opts_t *o);

/* Decoded input records (doc/binary_format.md), ids dense and in id order.  Borrowed
 * for the duration of the call only. */
typedef struct dwg_raw {
  uint64_t nvar, ndom, nfac, nweight;
  const uint8_t *var_role;   /* [nvar] 0 query, >=1 evidence (src/binary_format.cc:107) */
  const uint64_t *var_init;  /* [nvar] initial (sparse) value */
  const uint16_t *var_dtype; /* [nvar] 0 Boolean, 1 categorical */
  const uint64_t *var_card;  /* [nvar] */
  const uint64_t *dom_vid;   /* [ndom] */
  const uint64_t *dom_ptr;   /* [ndom+1] */
  const uint64_t *dom_val;   /* [dom_ptr[ndom]] sparse values, file order = dense order */
  const double *dom_truth;   /* [dom_ptr[ndom]] */
  const int32_t *fac_func;   /* [nfac] DWG_FUNC_* */
  const uint64_t *fac_ptr;   /* [nfac+1] CSR over edges */
  const uint64_t *edge_vid;  /* [fac_ptr[nfac]] */
  const uint64_t *edge_eq;   /* [fac_ptr[nfac]] ORIGINAL (sparse) equal_to */
  const uint64_t *fac_wid;   /* [nfac] */
  const double *fac_fv;      /* [nfac] */
  const double *w_init;      /* [nweight] */
  const uint8_t *w_fixed;    /* [nweight] */
} dwg_raw_t;

typedef struct dwg_graph dwg_graph;
typedef struct dwg_handle dwg_handle;

const char *dwg_last_error(void);
const char *dwg_version(void);

/* ---- graph: replaces FactorGraph::load_* (src/binary_format.cc:48-226),
 * safety_check (src/factor_graph.cc:201-219) and construct_index (:90-199) -------- */
int dwg_graph_from_raw(const dwg_raw_t *raw, int schedule, dwg_graph **out);
/* files may be pipes; each kind may have several files (src/cmd_parser.cc:63-70);
 * counts are checked against the meta file like safety_check does */
int dwg_graph_from_files(const char *meta, int n_var, const char *const *var_files, int n_dom,
                         const char *const *dom_files, int n_fac, const char *const *fac_files,
                         int n_wgt, const char *const *wgt_files, int schedule, dwg_graph **out);
/* Benchmark input tooling (no reference counterpart): the compiled image of a synthetic graph of
 * nvar Boolean variables x `degree` unary ISTRUE factors over nweight tied weights (ids uniform, or
 * Zipf(zipf) when zipf > 0; feature values 1 or ~N(0,1) when real_features; evidence with
 * probability p_evid, labelled by a planted logistic model), streamed slice by slice so that graphs
 * of 10^9 .. 3 x 10^9 factors (BASELINE.json configs[3], configs[4]) never exist as records.  The
 * same graph, record for record, comes out of sampler_b200/graphgen.py synth_unary.  cdf_out
 * (optional, nweight doubles): the Zipf table used. */
int dwg_graph_synthetic_unary(uint64_t nvar, uint32_t degree, uint64_t nweight, double zipf, int real_features,
                              double p_evid, uint64_t seed, double *cdf_out, dwg_graph **out);
void dwg_graph_free(dwg_graph *g);
/* The compiled image is page-locked on the first dwg_create so that every replica's upload is a
 * DMA at the PCIe rate (the reference deep-copies the FactorGraph per NUMA node,
 * src/dimmwitted.cc:113-116).  Page-locking costs about three pageable uploads: a host that
 * creates ONE replica and exits (dw gibbs -c 1) turns it off. */
int dwg_graph_set_page_locking(dwg_graph *g, int on);

typedef struct dwg_graph_info {
  uint64_t num_variables, num_factors, num_weights, num_edges, num_values;
  uint64_t num_variables_evidence, num_variables_query; /* FactorGraphDescriptor, factor_graph.h:18-46 */
  uint64_t num_index_entries;   /* |factor_index| after dedupe (factor_graph.cc:179-195) */
  uint64_t num_other_refs;      /* words in the others stream */
  uint32_t num_colors;          /* 1 = hogwild */
  int schedule;                 /* resolved DWG_SCHED_* */
  int has_categorical, has_multi_arity, fv_mode /*0 all 1.0, 1 f32, 2 f64*/, assign_bytes;
  uint64_t device_bytes;        /* graph + state bytes one replica occupies in HBM */
  /* ALGORITHMIC bytes of one sweep: inference over all variables
   * (--sample_evidence), over query variables only, and one learning sweep */
  double bytes_per_sweep_infer, bytes_per_sweep_infer_query, bytes_per_sweep_learn;
  uint64_t num_device_rows;     /* length of the device tally array (dwg_tallies_device_ptr) */
} dwg_graph_info_t;
int dwg_graph_info(const dwg_graph *g, dwg_graph_info_t *info);
/* per-variable metadata in original id order, for writers (src/inference_result.cc:211-243):
 * dtype[nvar], is_evid[nvar], value_base[nvar+1]; values[num_values] = sparse domain values */
int dwg_graph_var_meta(const dwg_graph *g, uint8_t *dtype, uint8_t *is_evid, uint64_t *value_base,
                       uint64_t *values);

/* ---- replica: replaces GibbsSampler ctor + InferenceResult ctor
 * (src/gibbs_sampler.cc:5-18, src/inference_result.cc:24-42) -------------------- */
int dwg_create(const dwg_graph *g, const dwg_opts_t *opts, int device, dwg_handle **out);
void dwg_destroy(dwg_handle *h);
/* Device memory of destroyed replicas stays in the device's stream-ordered pool for the next
 * dwg_create (no equivalent in the reference: its replicas live as long as the process).  This
 * returns what the pool holds unused to the driver. */
int dwg_release_cached_memory(int device);
/* The device arrays that hold the replica's copy of the graph (stream, per-variable tables, initial
 * weights), in a fixed order that depends on the graph only: ptrs[i] / bytes[i] for i < *n (call
 * with max = 0 to get the count).  What a replica created with defer_upload receives from a
 * peer (torch.distributed / NCCL broadcast on dwg_stream(), or cudaMemcpyPeer).  The map that
 * orders tallies for the host is not among them: a deferred replica that is asked for tallies
 * uploads its own. */
int dwg_graph_segments(dwg_handle *h, int max, void **ptrs, uint64_t *bytes, int *n);
/* InferenceResult ctor again (src/inference_result.cc:24-42): both chains to the evidence values,
 * tallies cleared; the weights stay as they are. */
int dwg_reset_state(dwg_handle *h);

/* GibbsSampler::sample_sgd(stepsize)+wait (src/gibbs_sampler.cc:27-38): one learning sweep.
 * Asynchronous on the handle's stream. */
int dwg_learn_epoch(dwg_handle *h, double stepsize);
/* GibbsSampler::sample(i_epoch)+wait (src/gibbs_sampler.cc:20-25): one inference sweep. */
int dwg_infer_epoch(dwg_handle *h);
/* n sweeps back to back (the launches of consecutive colour classes are chained by programmatic
 * dependent launch; no host synchronisation in between) */
int dwg_infer_epochs(dwg_handle *h, uint64_t n);
/* n learning sweeps of ONE replica with stepsize *= decay after each (DimmWitted::learn without the
 * per-epoch replica averaging, src/dimmwitted.cc:162-207).  Under the sequential schedule (tiny
 * graphs) dwg_infer_epochs / dwg_learn_epochs run all n epochs in one launch: such graphs are launch
 * latency and nothing else. */
int dwg_learn_epochs(dwg_handle *h, uint64_t n, double stepsize, double decay);
int dwg_sync(dwg_handle *h);
void *dwg_stream(dwg_handle *h); /* cudaStream_t the sweeps are launched on */

/* InferenceResult::clear_variabletally, src/inference_result.cc:107-114 */
int dwg_clear_tallies(dwg_handle *h);
/* InferenceResult state accessors (src/inference_result.h:27-43); synchronise first.
 * tallies[num_values], nsamples[num_variables] in original (variable, domain-index) order */
int dwg_get_tallies(dwg_handle *h, uint64_t *tallies, uint64_t *nsamples);
/* the same with the tallies at the device's own counter width (u32: no widening pass on the host;
 * a tally cannot exceed the number of sweeps).  Page-locked destinations (cudaHostAlloc /
 * cudaHostRegister memory) are filled by DMA straight from the device -- the tallies of
 * dwg_get_tallies32 and the sample counts of both getters; pageable ones by the host's threads.
 * Either pointer may be NULL. */
int dwg_get_tallies32(dwg_handle *h, uint32_t *tallies, uint64_t *nsamples);
/* aggregate_marginals_from :116-127.  The device counts in 32 bits: an aggregate of more than
 * 2^32 - 1 samples of one variable is refused (DWG_ERR_UNSUPPORTED), never wrapped */
int dwg_add_tallies(dwg_handle *h, const uint64_t *tallies, const uint64_t *nsamples);
int dwg_get_weights(dwg_handle *h, double *out);
int dwg_set_weights(dwg_handle *h, const double *in);
int dwg_get_assignments(dwg_handle *h, int evid_chain, uint64_t *out); /* dense values */
int dwg_set_assignments(dwg_handle *h, int evid_chain, const uint64_t *in);

/* ---- replica averaging: replaces DimmWitted::update_weights + copy_weights_to
 * (src/dimmwitted.cc:209-243,199-200; src/inference_result.cc:68-86) -------------
 * Device pointer to the f64[num_weights] weight vector, for an external allreduce
 * (torch.distributed / NCCL) on dwg_stream(). */
void *dwg_weights_device_ptr(dwg_handle *h);
/* After the weights were SUMMED over n_replicas: scale non-fixed weights by 1/n_replicas
 * (fixed ones are restored), compute lmax / l2 of the change since the previous call
 * (dimmwitted.cc:220-227) in one epilogue kernel.  lmax/l2 may be NULL. */
int dwg_finish_weight_average(dwg_handle *h, int n_replicas, double *lmax, double *l2);
/* The same exchange through HOST memory (no NCCL: e.g. replicas in different processes talking
 * gloo/MPI): store the cross-replica SUM of dwg_get_weights() vectors (original weight order),
 * then call dwg_finish_weight_average.  merge_weights_from, src/inference_result.cc:68-73. */
int dwg_set_weight_sum(dwg_handle *h, const double *sum);
/* Same, with the allreduce done here over an ncclComm_t (libnccl is dlopen'ed). */
int dwg_average_weights_nccl(dwg_handle *h, void *nccl_comm, int n_replicas, double *lmax, double *l2);
/* In-process replicas (the CLI's -c N: one host thread driving N GPUs, like the reference's
 * one process driving N NUMA replicas).  comms: n ncclComm_t from dwg_comm_init_all, or NULL
 * to exchange through the host (no NCCL available).  Averages the weight vectors of all
 * handles (dimmwitted.cc:214-216,199-200); norms as above, from replica 0. */
int dwg_comm_init_all(int n, const int *devices, void **comms_out); /* ncclCommInitAll */
int dwg_comm_destroy(void *comm);
int dwg_average_weights_multi(dwg_handle *const *hs, void *const *comms, int n, double *lmax, double *l2);
/* aggregate_marginals_from (inference_result.cc:116-127): sum tallies / sample counts of all
 * handles into hs[0] */
int dwg_reduce_tallies_multi(dwg_handle *const *hs, void *const *comms, int n);
/* Device pointers to u32 tallies[num_values] (schedule order) for an external reduce */
void *dwg_tallies_device_ptr(dwg_handle *h);
/* The same exchange with narrow counters (aggregate_marginals_from, src/inference_result.cc:116-127,
 * moves size_t per value; a tally never exceeds the sweeps summed over the replicas).  pack: `fields`
 * (1, 2 or 4) consecutive tallies of 32 / fields bits each per u32 word into a device buffer of
 * *nwords words, on the handle's stream -- the caller sums the words over the replicas (no carry
 * crosses a field as long as every aggregated tally fits its field) -- unpack: the words back into
 * the tally array.  10 M tallies of a 20-sweep x 8-replica run travel as 10 MB instead of 40 MB. */
int dwg_tallies_pack(dwg_handle *h, int fields, void **words, uint64_t *nwords);
int dwg_tallies_unpack(dwg_handle *h, int fields);

/* ---- parity hooks (tests only) ---------------------------------------------- */
/* FactorGraph::potential (src/factor_graph.h:127-145) for n (variable, dense proposal) pairs
 * on the handle's current chain and weights. */
int dwg_eval_potentials(dwg_handle *h, uint64_t n, const uint64_t *vids, const uint64_t *proposals,
                        int evid_chain, double *out);
/* draw_sample's decision given potentials and the uniform (src/gibbs_sampler.h:198-246):
 * n draws; pots row-major [n][ld]; card[i]==0 means Boolean with pots[i][0]=pos, [1]=neg. */
int dwg_draw(int device, uint64_t n, const double *pots, uint64_t ld, const uint32_t *card,
             const double *r, uint32_t *out);
/* logadd, src/common.h:118-132 */
int dwg_logadd(int device, uint64_t n, const double *a, const double *b, double *out);
/* Replace the counter RNG by explicit uniforms (original id order; NULL restores the RNG). */
int dwg_debug_set_uniforms(dwg_handle *h, const double *r_free, const double *r_evid);
/* Run sweeps with ONE thread in schedule order: the reference's -t 1 semantics exactly. */
int dwg_debug_set_sequential(dwg_handle *h, int on);
/* schedule order: perm[d] = original vid; color_start[num_colors+1] */
int dwg_debug_schedule(const dwg_graph *g, uint64_t *perm, uint64_t *color_start);
/* the order in which a LEARNING sweep of the parallel schedules visits the variables (slices in
 * ascending id block; the sequential / CSR schedules: schedule order); perm[num_variables] */
int dwg_debug_learn_order(const dwg_graph *g, uint64_t *perm, uint64_t *n_out);
/* borrow one array of the compiled image by name ("row_ptr", "e_wid", "e_tag", "oth", ...) so
 * CPU tests can check the device layout without a device; *elem_size in bytes */
int dwg_debug_image(const dwg_graph *g, const char *name, const void **data, uint64_t *count,
                    uint32_t *elem_size);
/* count of kernels this library launched in this process (bench.py's gpu_launches) */
uint64_t dwg_launch_count(void);

/* ---- the whole CLI: replaces dd::dw (src/dimmwitted.cc:20-35) ------------------ */
int dwg_dw_main(int argc, const char *const argv[]);

#ifdef __cplusplus
}
#endif
#endif /* DWG_H_ */
