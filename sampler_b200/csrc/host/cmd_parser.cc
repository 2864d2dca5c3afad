// cmd_parser.cc -- see cmd_parser.h.  Behaviour kept from the reference
// (src/cmd_parser.cc): every gibbs flag with its short and long name (:59-113), repeated
// flags take the LAST value (:18-21) and file flags accumulate, -l and -i are required,
// unknown flags are an error (TCLAP prints the problem and exits 1), the -a/-p alias hack
// (:156-160), defaults decay 0.95 / reg_param 0.01 / l2 (:161-166).  -c counts GPUs here,
// not NUMA nodes: 0 (default) = every visible GPU, like "all NUMA nodes" (:127-136).
#include "cmd_parser.h"

#include <cstdlib>
#include <iostream>
#include <map>

namespace dwg {
namespace {

struct Spec {
  const char *shrt, *lng;
  int kind;  // 0 switch, 1 value
};
const Spec GIBBS_SPECS[] = {
    {"m", "fg_meta", 1}, {"v", "variables", 1}, {"", "domains", 1}, {"f", "factors", 1},
    {"w", "weights", 1}, {"o", "outputFile", 1}, {"l", "n_learning_epoch", 1},
    {"i", "n_inference_epoch", 1}, {"", "burn_in", 1}, {"c", "n_datacopy", 1},
    {"t", "n_threads", 1}, {"a", "alpha", 1}, {"p", "stepsize", 1}, {"d", "diminish", 1},
    {"b", "reg_param", 1}, {"", "regularization", 1}, {"q", "quiet", 0},
    {"", "sample_evidence", 0}, {"", "learn_non_evidence", 0}, {"", "noise_aware", 0},
    {"", "schedule", 1}, {"", "seed", 1}, {"", "learn_chunks", 1},
};
const Spec BIN2TEXT_SPECS[] = {
    {"m", "fg_meta", 1}, {"v", "variables", 1}, {"", "domains", 1}, {"f", "factors", 1},
    {"w", "weights", 1}, {"o", "outputFile", 1},
};

bool parse_u64(const std::string &s, uint64_t *out) {
  if (s.empty() || s[0] == '-') return false;
  char *end = nullptr;
  unsigned long long v = std::strtoull(s.c_str(), &end, 10);
  if (!end || *end) return false;
  *out = v;
  return true;
}
bool parse_f64(const std::string &s, double *out) {
  if (s.empty()) return false;
  char *end = nullptr;
  double v = std::strtod(s.c_str(), &end);
  if (!end || *end) return false;
  *out = v;
  return true;
}

void usage(const char *arg0) {
  std::cout << "DimmWitted (https://github.com/HazyResearch/sampler) -- B200-native Gibbs path" << std::endl;
  std::cout << "Usage: " << arg0 << " MODE [ARG...]" << std::endl;
  for (const char *m : {"bin2text", "gibbs", "text2bin"}) std::cout << "  " << arg0 << " " << m << std::endl;
}

}  // namespace

CmdParser::CmdParser(int argc, const char *const argv[], int num_gpus) {
  const char *arg0 = argc > 0 ? argv[0] : "dw";
  app_name = argc > 1 ? argv[1] : "";
  auto error = [&](const std::string &msg) {
    std::cerr << "PARSE ERROR: " << msg << std::endl;
    std::cerr << "For complete USAGE and HELP type: " << arg0 << " " << app_name << " --help" << std::endl;
    ++num_errors;
  };

  if (app_name == "gibbs" || app_name == "bin2text") {
    const Spec *specs = app_name == "gibbs" ? GIBBS_SPECS : BIN2TEXT_SPECS;
    size_t nspec = app_name == "gibbs" ? sizeof(GIBBS_SPECS) / sizeof(Spec)
                                       : sizeof(BIN2TEXT_SPECS) / sizeof(Spec);
    std::map<std::string, std::vector<std::string>> vals;  // by long name
    std::map<std::string, int> switches;
    for (int i = 2; i < argc; ++i) {
      std::string a = argv[i];
      if (a == "--") break;  // TCLAP's ignore-rest marker
      if (a == "-h" || a == "--help" || a == "--version") {
        if (a == "--version") std::cout << arg0 << " " << app_name << "  version: 0.01" << std::endl;
        else {
          std::cout << "USAGE: " << arg0 << " " << app_name;
          for (size_t k = 0; k < nspec; ++k)
            std::cout << " [--" << specs[k].lng << (specs[k].kind ? " <value>]" : "]");
          std::cout << std::endl;
        }
        exit_now = true;
        return;
      }
      const Spec *hit = nullptr;
      if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
        for (size_t k = 0; k < nspec; ++k)
          if (a.substr(2) == specs[k].lng) hit = &specs[k];
      } else if (a.size() == 2 && a[0] == '-') {
        for (size_t k = 0; k < nspec; ++k)
          if (specs[k].shrt[0] && a.substr(1) == specs[k].shrt) hit = &specs[k];
      }
      if (!hit) {
        error("Argument: " + a + "\n             Couldn't find match for argument");
        continue;
      }
      if (hit->kind == 0) {
        ++switches[hit->lng];
      } else {
        if (i + 1 >= argc) {
          error(std::string("Argument: --") + hit->lng + "\n             Missing a value for this argument!");
          break;
        }
        vals[hit->lng].push_back(argv[++i]);
      }
    }
    auto last = [&](const char *name, const std::string &dflt) {
      auto it = vals.find(name);
      return it == vals.end() ? dflt : it->second.back();
    };
    auto last_u64 = [&](const char *name, uint64_t dflt) {
      auto it = vals.find(name);
      if (it == vals.end()) return dflt;
      uint64_t best = dflt;
      for (auto &s : it->second)
        if (!parse_u64(s, &best)) error(std::string("Argument: --") + name + "\n             Couldn't read argument value from string '" + s + "'");
      return best;
    };
    auto last_f64 = [&](const char *name, double dflt) {
      auto it = vals.find(name);
      if (it == vals.end()) return dflt;
      double best = dflt;
      for (auto &s : it->second)
        if (!parse_f64(s, &best)) error(std::string("Argument: --") + name + "\n             Couldn't read argument value from string '" + s + "'");
      return best;
    };
    fg_file = last("fg_meta", "");
    variable_file = vals["variables"];
    domain_file = vals["domains"];
    factor_file = vals["factors"];
    weight_file = vals["weights"];
    output_folder = last("outputFile", "");
    if (app_name == "bin2text") return;

    for (const char *req : {"n_learning_epoch", "n_inference_epoch"})
      if (!vals.count(req)) error(std::string("Required argument missing: ") + req);
    n_learning_epoch = last_u64("n_learning_epoch", 0);
    n_inference_epoch = last_u64("n_inference_epoch", 0);

    n_datacopy = last_u64("n_datacopy", 0);
    const uint64_t avail = num_gpus > 0 ? (uint64_t)num_gpus : 1;
    if (n_datacopy == 0) n_datacopy = avail;
    if (!(0 < n_datacopy && n_datacopy <= avail)) {
      std::cerr << "n_datacopy (" << n_datacopy << ") must be in the range of [0, " << avail
                << "] (number of GPUs)" << std::endl;
      ++num_errors;
    }
    n_threads = last_u64("n_threads", 0);  // host threads do not sample here; kept and echoed

    burn_in = last_u64("burn_in", 0);
    stepsize = last_f64("alpha", 0.01);
    stepsize2 = last_f64("stepsize", 0.01);
    if (stepsize == 0.01) stepsize = stepsize2;  // the reference's two-flag hack, :156-160
    decay = last_f64("diminish", 0.95);
    reg_param = last_f64("reg_param", 0.01);
    reg_l1 = last("regularization", "l2") == "l1";

    should_be_quiet = switches["quiet"] > 0;
    should_sample_evidence = switches["sample_evidence"] > 0;
    should_learn_non_evidence = switches["learn_non_evidence"] > 0;
    is_noise_aware = switches["noise_aware"] > 0;

    std::string sch = last("schedule", "auto");
    if (sch == "auto") schedule = 0;
    else if (sch == "hogwild") schedule = 1;
    else if (sch == "chromatic") schedule = 2;
    else if (sch == "sequential") schedule = 3;
    else error("Argument: --schedule\n             expects auto, hogwild, chromatic or sequential");
    seed_given = vals.count("seed") > 0;
    seed = last_u64("seed", seed);
    learn_chunks = (int)last_u64("learn_chunks", 0);

  } else if (app_name == "text2bin") {
    // dw text2bin MODE INPUT OUTPUT COUNT_OUTPUT [FUNC_ID ARITY VAR_IS_POSITIVE...]  (:173-219)
    std::vector<std::string> pos;
    for (int i = 2; i < argc; ++i) pos.push_back(argv[i]);
    if (pos.size() < 4) {
      error("Required arguments missing: mode, input, output, count_output");
      return;
    }
    text2bin_mode = pos[0];
    text2bin_input = pos[1];
    text2bin_output = pos[2];
    text2bin_count_output = pos[3];
    if (text2bin_mode == "factor") {
      if (pos.size() < 7) {
        error("Required arguments missing: func_id, arity, var_is_positive");
        return;
      }
      text2bin_factor_func_id = std::atoi(pos[4].c_str());
      uint64_t a = 0;
      if (!parse_u64(pos[5], &a)) error("arity: not a number");
      text2bin_factor_arity = a;
      for (size_t i = 6; i < pos.size(); ++i) {
        uint64_t v = 0;
        if (!parse_u64(pos[i], &v)) error("var_is_positive: not a number");
        text2bin_factor_variables_should_equal_to.push_back(v);
      }
    } else if (pos.size() > 4) {
      error("Argument: " + pos[4] + "\n             Couldn't find match for argument");
    }
  } else {
    // unknown or missing mode (:235-246)
    ++num_errors;
    usage(arg0);
    if (argc > 1) std::cerr << app_name << ": Unrecognized MODE" << std::endl;
  }
}

static std::ostream &operator<<(std::ostream &out, const std::vector<std::string> &v) {
  // "[a, b, \b\b]" exactly as the reference's vector printer does (:260-268)
  if (!v.empty()) {
    out << '[';
    for (auto &s : v) out << s << ", ";
    out << "\b\b]";
  }
  return out;
}

std::ostream &operator<<(std::ostream &stream, const CmdParser &args) {  // :271-292
  stream << "#################GIBBS SAMPLING#################" << std::endl;
  stream << "# fg_file            : " << args.fg_file << std::endl;
  stream << "# variable_file      : " << args.variable_file << std::endl;
  stream << "# domain_file        : " << args.domain_file << std::endl;
  stream << "# weight_file        : " << args.weight_file << std::endl;
  stream << "# factor_file        : " << args.factor_file << std::endl;
  stream << "# output_folder      : " << args.output_folder << std::endl;
  stream << "# n_learning_epoch   : " << args.n_learning_epoch << std::endl;
  stream << "# n_inference_epoch  : " << args.n_inference_epoch << std::endl;
  stream << "# stepsize           : " << args.stepsize << std::endl;
  stream << "# decay              : " << args.decay << std::endl;
  stream << "# regularization     : " << args.reg_param << std::endl;
  stream << "# burn_in            : " << args.burn_in << std::endl;
  stream << "# n_datacopy         : " << args.n_datacopy << std::endl;
  stream << "# n_threads          : " << args.n_threads << std::endl;
  stream << "# learn_non_evidence : " << args.should_learn_non_evidence << std::endl;
  stream << "# is_noise_aware     : " << args.is_noise_aware << std::endl;
  stream << "################################################" << std::endl;
  return stream;
}

}  // namespace dwg
