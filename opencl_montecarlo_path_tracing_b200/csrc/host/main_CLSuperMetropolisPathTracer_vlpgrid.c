/* drop-in for CLSuperMetropolisPathTracer_vlpgrid/CLSuperMetropolisPathTracer.c (FIX mode: see pthost_cli_metro.c) */
#include "pthost.h"
int main(int argc, char **argv) { return pth_cli_metropolis_main(argc, argv); }
