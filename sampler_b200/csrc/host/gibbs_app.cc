// gibbs_app.cc -- `dw` front end: dispatch (dd::dw, src/dimmwitted.cc:20-35) and the
// `dw gibbs` pipeline + learn/inference schedule (dd::gibbs :37-95, DimmWitted::learn
// :162-207, update_weights :209-243, inference :121-160, dump_weights :245-258,
// aggregate_results_and_dump :260-277, compute_n_epochs :280-282) and the text writers
// (InferenceResult::dump_weights_in_text / dump_marginals_in_text / snippets / histogram,
// src/inference_result.cc:88-243).  All device work goes through the C ABI (include/dwg.h):
// "NUMA node" became "GPU", one replica per GPU (-c), driven from this one host thread.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime_api.h>

#include "../../../include/dwg.h"
#include "cmd_parser.h"
#include "parallel.h"

namespace dwg {
int text2bin_main(const CmdParser &args);
int bin2text_main(const CmdParser &args);

namespace {

// Text dump of n items: ranges are formatted on the host threads into their own buffers and the
// buffers written in order -- the file is byte for byte the sequential one (dump_*_in_text,
// src/inference_result.cc:101-105,211-243, one ostream << per field there).
template <typename Fmt>
bool dump_lines(const std::string &fn, uint64_t n, Fmt fmt) {
  FILE *f = std::fopen(fn.c_str(), "w");
  if (!f) return false;
  const unsigned parts = n >= (1u << 16) ? host_threads() * 4 : 1;
  std::vector<std::string> buf(parts);
  parallel_ranges(n, parts, [&](uint64_t a, uint64_t b, unsigned p) {
    std::string &out = buf[p];
    out.reserve((b - a) * 24);
    char line[96];
    for (uint64_t i = a; i < b; ++i) fmt(i, out, line);
  });
  bool ok = true;
  for (const std::string &b : buf) ok = ok && std::fwrite(b.data(), 1, b.size(), f) == b.size();
  return std::fclose(f) == 0 && ok;
}

struct Timer {  // src/timer.cc:21-28 (CLOCK_MONOTONIC stopwatch, float seconds)
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void restart() { t0 = std::chrono::steady_clock::now(); }
  float elapsed() const {
    return std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count();
  }
};

// FactorGraphDescriptor printer, src/factor_graph.cc:25-44
void print_desc(std::ostream &s, uint64_t nvar, uint64_t nqry, uint64_t nevd, uint64_t nfac,
                uint64_t nw, uint64_t ne, uint64_t nval) {
  s << "#V=" << nvar;
  if (nqry + nevd > 0) s << "(#Vqry=" << nqry << " #Vevd=" << nevd << ")";
  s << " #F=" << nfac << " #W=" << nw << " #E=" << ne << " #Val=" << nval;
}

#define TRY(call)                                                          \
  do {                                                                     \
    if ((call) != DWG_OK) {                                                \
      std::cerr << "[ERROR] " << dwg_last_error() << std::endl;            \
      return 1;                                                            \
    }                                                                      \
  } while (0)

struct Replicas {
  std::vector<dwg_handle *> hs;
  std::vector<void *> comms;
  dwg_graph *graph = nullptr;
  ~Replicas() {
    for (auto *h : hs) dwg_destroy(h);
    for (auto *c : comms) dwg_comm_destroy(c);
    dwg_graph_free(graph);
  }
};

int gibbs(const CmdParser &args) {
  int ngpu = 0;
  if (cudaGetDeviceCount(&ngpu) != cudaSuccess) ngpu = 0;
  if (!args.should_be_quiet) {
    std::cout << std::endl;
    std::cout << "#################MACHINE CONFIG#################" << std::endl;
    std::cout << "# # GPU (replica slots): " << ngpu << std::endl;
    std::cout << "################################################" << std::endl;
    std::cout << std::endl;
    std::cout << args << std::endl;
  }
  if (ngpu == 0) {
    std::cerr << "[ERROR] no CUDA device: the Gibbs sweep has no CPU path" << std::endl;
    return 1;
  }

  Replicas R;
  // DWG_COMPILE_TIMING=1: wall time of the application's phases on stderr, like the loader's and compiler's
  const bool timing = getenv("DWG_COMPILE_TIMING") != nullptr;
  Timer t_phase;
  auto lap = [&](const char *what) {
    if (timing) std::fprintf(stderr, "[app]     %-28s %8.1f ms\n", what, 1e3 * t_phase.elapsed());
    t_phase.restart();
  };
  // the CUDA context comes up (0.3-0.5 s) on a second thread while the graph is loaded and compiled
  std::thread cuda_up([] { cudaFree(nullptr); });
  struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } cuda_up_joiner{cuda_up};
  std::cout << "\tloading factor graph..." << std::endl;
  auto cptr = [](const std::vector<std::string> &v) {
    std::vector<const char *> p;
    for (auto &s : v) p.push_back(s.c_str());
    if (p.empty()) p.push_back(nullptr);
    return p;
  };
  auto vf = cptr(args.variable_file), df = cptr(args.domain_file), ff = cptr(args.factor_file),
       wf = cptr(args.weight_file);
  TRY(dwg_graph_from_files(args.fg_file.c_str(), (int)args.variable_file.size(), vf.data(),
                           (int)args.domain_file.size(), df.data(), (int)args.factor_file.size(),
                           ff.data(), (int)args.weight_file.size(), wf.data(), args.schedule, &R.graph));
  lap("load + compile");
  dwg_graph_info_t info;
  TRY(dwg_graph_info(R.graph, &info));
  std::cout << "Factor graph loaded:\t";
  print_desc(std::cout, info.num_variables, info.num_variables_query, info.num_variables_evidence,
             info.num_factors, info.num_weights, info.num_edges, 0);
  std::cout << std::endl << "Factor graph indexed:\t";
  print_desc(std::cout, info.num_variables, info.num_variables_query, info.num_variables_evidence,
             info.num_factors, info.num_weights, info.num_edges, info.num_values);
  std::cout << std::endl;
  if (!args.should_be_quiet) {
    std::cout << "Device layout: " << info.num_index_entries << " edge records, "
              << info.num_other_refs << " other-variable refs, schedule="
              << (info.schedule == DWG_SCHED_CHROMATIC ? "chromatic"
                  : info.schedule == DWG_SCHED_SEQUENTIAL ? "sequential" : "hogwild") << " ("
              << info.num_colors << " colour" << (info.num_colors == 1 ? "" : "s") << "), "
              << info.device_bytes / 1048576.0 << " MiB per replica" << std::endl;
  }

  // DimmWitted ctor, src/dimmwitted.cc:97-119: one replica per GPU
  const size_t n_samplers = args.n_datacopy;
  dwg_opts_t opts;
  dwg_opts_default(&opts);
  opts.sample_evidence = args.should_sample_evidence;
  opts.learn_non_evidence = args.should_learn_non_evidence;
  opts.noise_aware = args.is_noise_aware;
  opts.regularization = args.reg_l1 ? 1 : 0;
  opts.reg_param = args.reg_param;
  opts.schedule = args.schedule;
  // the reference seeds from un-seeded rand() (gibbs_sampler.cc:49): a fixed default; --seed varies it
  opts.seed = args.seed;
  opts.learn_chunks = args.learn_chunks;
  if (n_samplers == 1) TRY(dwg_graph_set_page_locking(R.graph, 0));  // one upload: not worth locking 0.6 GB
  for (size_t i = 0; i < n_samplers; ++i) {
    std::cout << "CREATE CFG ON GPU ... " << i << std::endl;
    opts.replica = (uint32_t)i;
    dwg_handle *h = nullptr;
    TRY(dwg_create(R.graph, &opts, (int)i, &h));
    R.hs.push_back(h);
  }
  if (n_samplers > 1) {
    std::vector<int> devs(n_samplers);
    for (size_t i = 0; i < n_samplers; ++i) devs[i] = (int)i;
    R.comms.resize(n_samplers, nullptr);
    if (dwg_comm_init_all((int)n_samplers, devs.data(), R.comms.data()) != DWG_OK) {
      std::cerr << "[WARN] NCCL unavailable (" << dwg_last_error()
                << "); averaging replicas through the host" << std::endl;
      R.comms.clear();
    }
  }
  lap("replicas up (CUDA context, upload)");
  void *const *comms = R.comms.empty() ? nullptr : R.comms.data();
  auto compute_n_epochs = [&](uint64_t n) {  // :280-282
    return (uint64_t)std::ceil((double)n / n_samplers);
  };
  const uint64_t nvar = info.num_variables, nweight = info.num_weights;
  const bool progress = !args.should_be_quiet;

  // ---- learn, :162-207 ----
  {
    const uint64_t n_epoch = compute_n_epochs(args.n_learning_epoch);
    Timer t_total, t;
    double stepsize = args.stepsize;
    // one quiet replica: no per-epoch averaging, no per-epoch line -- all epochs in one call (one
    // launch under the sequential schedule of tiny graphs, which are launch latency and nothing else)
    const bool batched = !progress && n_samplers == 1;
    if (batched) TRY(dwg_learn_epochs(R.hs[0], n_epoch, stepsize, args.decay));
    for (uint64_t e = 0; e < n_epoch && !batched; ++e) {
      if (progress) {
        std::cout << std::setprecision(3) << "LEARNING EPOCH " << e * n_samplers << "~"
                  << ((e + 1) * n_samplers - 1) << "...." << std::flush << std::setprecision(6);
      }
      t.restart();
      for (auto *h : R.hs) TRY(dwg_learn_epoch(h, stepsize));
      if (progress || n_samplers > 1) {
        double lmax = 0, l2 = 0;
        TRY(dwg_average_weights_multi(R.hs.data(), comms, (int)n_samplers, progress ? &lmax : nullptr,
                                      progress ? &l2 : nullptr));
        if (progress) {
          for (auto *h : R.hs) TRY(dwg_sync(h));
          double el = t.elapsed();
          std::cout << std::setprecision(3) << "" << el << " sec.," << (nvar * n_samplers) / el
                    << " vars/sec.,stepsize=" << stepsize << ",lmax=" << lmax / stepsize
                    << ",l2=" << std::sqrt(l2) / stepsize << std::endl << std::setprecision(6);
        }
      }
      stepsize *= args.decay;  // :202
    }
    for (auto *h : R.hs) TRY(dwg_sync(h));
    std::cout << "TOTAL LEARNING TIME: " << t_total.elapsed() << " sec." << std::endl;
    lap("learning");
  }

  // ---- dump_weights, :245-258 (written even with -l 0) ----
  std::vector<double> w(nweight);
  TRY(dwg_get_weights(R.hs[0], w.data()));
  if (progress) {
    std::cout << "LEARNING SNIPPETS (QUERY WEIGHTS):" << std::endl;  // inference_result.cc:88-99
    for (uint64_t j = 0; j < nweight && j < 10; ++j) std::cout << "   " << j << " " << w[j] << std::endl;
    std::cout << "   ..." << std::endl;
  }
  {
    std::string fn = args.output_folder + "/inference_result.out.weights.text";
    std::cout << "DUMPING... TEXT    : " << fn << std::endl;
    if (!dump_lines(fn, nweight, [&](uint64_t j, std::string &out, char *line) {
          out.append(line, (size_t)std::snprintf(line, 96, "%llu %g\n", (unsigned long long)j, w[j]));
        })) {
      std::cerr << "[ERROR] cannot write " << fn << std::endl;
      return 1;
    }
  }

  // ---- inference, :121-160.  --burn_in is parsed, printed and NOT used, like the reference
  // (src/cmd_parser.cc:91, no reader): identical arguments give identical schedules. ----
  {
    const uint64_t n_epoch = compute_n_epochs(args.n_inference_epoch);
    Timer t_total, t;
    for (auto *h : R.hs) TRY(dwg_clear_tallies(h));
    const bool batched = !progress;
    if (batched)
      for (auto *h : R.hs) TRY(dwg_infer_epochs(h, n_epoch));
    for (uint64_t e = 0; e < n_epoch && !batched; ++e) {
      if (progress) {
        std::cout << std::setprecision(3) << "INFERENCE EPOCH " << e * n_samplers << "~"
                  << ((e + 1) * n_samplers - 1) << "...." << std::flush << std::setprecision(6);
      }
      t.restart();
      for (auto *h : R.hs) TRY(dwg_infer_epoch(h));
      if (progress) {
        for (auto *h : R.hs) TRY(dwg_sync(h));
        double el = t.elapsed();
        std::cout << std::setprecision(3) << "" << el << " sec.," << (nvar * n_samplers) / el
                  << " vars/sec" << std::endl << std::setprecision(6);
      }
    }
    for (auto *h : R.hs) TRY(dwg_sync(h));
    std::cout << "TOTAL INFERENCE TIME: " << t_total.elapsed() << " sec." << std::endl;
    lap("weights dump + inference");
  }

  if (args.n_inference_epoch > 0) {  // :89-92
    // aggregate_results_and_dump, :260-277
    if (n_samplers > 1) TRY(dwg_reduce_tallies_multi(R.hs.data(), comms, (int)n_samplers));
    std::vector<uint64_t> tallies(info.num_values), nsamples(nvar), value_base(nvar + 1),
        values(info.num_values);
    std::vector<uint8_t> dtype(nvar), is_evid(nvar);
    TRY(dwg_get_tallies(R.hs[0], tallies.data(), nsamples.data()));
    TRY(dwg_graph_var_meta(R.graph, dtype.data(), is_evid.data(), value_base.data(), values.data()));
    auto prob = [&](uint64_t v, uint64_t k) {
      return 1.0 * tallies[value_base[v] + k] / nsamples[v];
    };
    if (progress) {  // show_marginal_snippet, inference_result.cc:129-169
      std::cout << "INFERENCE SNIPPETS (QUERY VARIABLES):" << std::endl;
      size_t ct = 0;
      for (uint64_t v = 0; v < nvar; ++v) {
        if (is_evid[v] && !args.should_sample_evidence) continue;
        ++ct;
        std::cout << "   " << v << "  NSAMPLE=" << nsamples[v] << std::endl;
        if (!dtype[v]) std::cout << "      @ " << 1 << " -> EXP=" << prob(v, 0) << std::endl;
        else
          for (uint64_t k = 0; k < value_base[v + 1] - value_base[v]; ++k)
            std::cout << "      @ " << values[value_base[v] + k] << " -> EXP=" << prob(v, k) << std::endl;
        if (ct % 10 == 0) break;
      }
      std::cout << "   ..." << std::endl;
    }
    std::string fn = args.output_folder + "/inference_result.out.text";
    std::cout << "DUMPING... TEXT    : " << fn << std::endl;
    // dump_marginals_in_text, inference_result.cc:211-243
    if (!dump_lines(fn, nvar, [&](uint64_t v, std::string &out, char *line) {
          if (is_evid[v] && !args.should_sample_evidence) return;
          if (!dtype[v]) {
            out.append(line, (size_t)std::snprintf(line, 96, "%llu 1 %g\n", (unsigned long long)v, prob(v, 0)));
          } else {
            for (uint64_t k = 0; k < value_base[v + 1] - value_base[v]; ++k)
              out.append(line, (size_t)std::snprintf(line, 96, "%llu %llu %g\n", (unsigned long long)v,
                                                     (unsigned long long)values[value_base[v] + k], prob(v, k)));
          }
        })) {
      std::cerr << "[ERROR] cannot write " << fn << std::endl;
      return 1;
    }
    if (progress) {  // show_marginal_histogram, inference_result.cc:171-209
      const size_t bins = 10;
      std::cout << "INFERENCE CALIBRATION (QUERY BINS):" << std::endl;
      std::vector<size_t> abc(bins + 1, 0);
      for (uint64_t v = 0; v < nvar; ++v) {
        if (!args.should_sample_evidence && is_evid[v]) continue;
        for (uint64_t k = 0; k < value_base[v + 1] - value_base[v]; ++k) {
          size_t bin = (size_t)((double)tallies[value_base[v] + k] / nsamples[v] * bins);
          if (bin <= bins) ++abc[bin];
        }
      }
      abc[bins - 1] += abc[bins];
      for (size_t i = 0; i < bins; ++i)
        std::printf("PROB BIN %.1f~%.1f  -->  # %zu\n", (float)i / bins, (float)(i + 1) / bins, abc[i]);
    }
  }
  lap("tallies + marginal dump");
  return 0;
}

}  // namespace
}  // namespace dwg

extern "C" int dwg_dw_main(int argc, const char *const argv[]) {
  using namespace dwg;
  int ngpu = 0;
  if (cudaGetDeviceCount(&ngpu) != cudaSuccess) ngpu = 0;
  CmdParser args(argc, argv, ngpu);
  if (args.exit_now) return args.exit_code;
  if (args.num_errors > 0) return args.num_errors;  // src/dimmwitted.cc:30
  try {
    if (args.app_name == "gibbs") return gibbs(args);
    if (args.app_name == "text2bin") return text2bin_main(args);
    if (args.app_name == "bin2text") return bin2text_main(args);
  } catch (const std::exception &e) {
    std::cerr << "[ERROR] " << e.what() << std::endl;
    return 1;
  }
  return 1;
}
