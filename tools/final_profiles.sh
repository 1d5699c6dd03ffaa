# round-2 final ncu captures (GPU box): the big-grid megakernel on the config-4 scene (3rd launch = longest-tile-first order),
# PT_KERNEL_SPEC's two passes on the config-1 scene, and the launch list of a short bench.py run
set -x
PT_DEAD_RAYS=elide QB_KERNELS=auto python tools/quick_bench.py base lmem 2>&1 | grep smem
PS_SPP=16 PS_REPS=3 PS_KERNEL=mega ncu --set full --import-source on --clock-control none -k regex:'k_mega_pixel' --launch-skip 2 --launch-count 1 -o gpurun_out/r2_25_soup_mega_final python tools/prof_soup.py > gpurun_out/r2_25_soup.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'k_spec' --launch-skip 4 --launch-count 2 -o gpurun_out/r2_26_spec_base_final python tools/profile_spec.py > gpurun_out/r2_26_spec.log 2>&1
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2_27_bench_short.json 2> gpurun_out/r2_27_bench_short.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_27_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2_27_bench_under_ncu.log 2>&1
tail -2 gpurun_out/r2_25_soup.log gpurun_out/r2_26_spec.log
