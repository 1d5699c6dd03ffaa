# quick timing of the headline kernels (GPU box): soup 16 spp (trace / elide), default scenes AUTO, config 3 at 128 spp
for dead in trace elide; do QG_KERNELS=mega QG_DEAD=$dead python tools/quick_grid.py 2>&1 | tail -1; done
for dead in trace elide; do echo "PT_DEAD_RAYS=$dead"; PT_DEAD_RAYS=$dead QB_KERNELS=auto python tools/quick_bench.py base lmem grid nodof 2>&1 | grep -v smem; done
for dead in trace elide; do echo "PT_DEAD_RAYS=$dead"; PT_DEAD_RAYS=$dead QC_SPP=128 python tools/quick_c3.py 2>&1 | grep "launch 3"; done
