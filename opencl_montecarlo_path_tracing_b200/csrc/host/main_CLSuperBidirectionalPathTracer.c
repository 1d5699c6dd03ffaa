/* Drop-in for the reference program CLSuperBidirectionalPathTracer/CLSuperBidirectionalPathTracer
 * (same argv incl. [N_VLP_per_light], scene files, stdout, result.ppm). */
#include "pthost.h"
int main(int argc, char **argv) { return pth_cli_main(PT_VARIANT_BIDIR, argc, argv); }
