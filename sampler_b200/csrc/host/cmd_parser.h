// cmd_parser.h -- the `dw` command line, flag for flag as the reference defines it
// (src/cmd_parser.h:16-76, src/cmd_parser.cc:56-247), without TCLAP.
#ifndef DWG_CMD_PARSER_H_
#define DWG_CMD_PARSER_H_
#include <cstdint>
#include <ostream>
#include <string>
#include <vector>

namespace dwg {

struct CmdParser {
  std::string app_name;  // gibbs | text2bin | bin2text

  // gibbs / bin2text
  std::string fg_file;
  std::vector<std::string> variable_file, domain_file, factor_file, weight_file;
  std::string output_folder;
  uint64_t n_learning_epoch = 0, n_inference_epoch = 0, n_datacopy = 0, n_threads = 0, burn_in = 0;
  double stepsize = 0.01, stepsize2 = 0.01, decay = 0.95, reg_param = 0.01;
  bool reg_l1 = false;
  bool should_be_quiet = false, should_sample_evidence = false, should_learn_non_evidence = false,
       is_noise_aware = false;
  // device-side additions (not in the reference)
  int schedule = 0;          // --schedule auto|hogwild|chromatic
  uint64_t seed = 0x5eed5eedull;  // --seed
  bool seed_given = false;
  int learn_chunks = 0;      // --learn_chunks N: weight updates per colour class and learning sweep (0 = auto)

  // text2bin
  std::string text2bin_mode, text2bin_input, text2bin_output, text2bin_count_output;
  int text2bin_factor_func_id = 0;
  uint64_t text2bin_factor_arity = 1;
  std::vector<uint64_t> text2bin_factor_variables_should_equal_to;

  int num_errors = 0;
  bool exit_now = false;  // --help / --version: usage printed, exit 0
  int exit_code = 0;

  // parse argv (argv[0] = program, argv[1] = mode); never exits the process
  CmdParser(int argc, const char *const argv[], int num_gpus);
};

std::ostream &operator<<(std::ostream &, const CmdParser &);

}  // namespace dwg
#endif
