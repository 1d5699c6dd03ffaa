/* Drop-in for the reference program CLSuperPathTracer_trianglegrid/CLSuperPathTracer (same argv, scene files, stdout, result.ppm). */
#include "pthost.h"
int main(int argc, char **argv) { return pth_cli_main(PT_VARIANT_GRID, argc, argv); }
