// dw_main.cc -- the `dw` executable: a main that delegates to the library, as the
// reference's src/main.cc:4 delegates to dd::dw.
#include "../../../include/dwg.h"
int main(int argc, char *argv[]) { return dwg_dw_main(argc, argv); }
