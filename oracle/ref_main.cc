// Test-infrastructure (NOT product code).  Seedable entry point for the
// UNMODIFIED reference: the reference seeds its per-thread erand48 state from
// un-seeded rand() (/root/reference/src/gibbs_sampler.cc:49) and has no seed
// flag, so "averaged over seeds" needs srand() before dd::dw()
// (/root/reference/src/dimmwitted.h:14, src/main.cc:4).
//   DW_SEED=<n> dw gibbs ...      (unset => identical to the stock binary)
#include <cstdlib>
namespace dd { int dw(int argc, const char *const argv[]); }
int main(int argc, char *argv[]) {
  if (const char *s = std::getenv("DW_SEED")) std::srand((unsigned)std::atol(s));
  return dd::dw(argc, argv);
}
