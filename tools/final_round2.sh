# end-of-round measurements on one B200: full GPU test suite, both bench arms, ncu capture of the headline kernel (text summary only)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_37_bench_n1.json 2> gpurun_out/r2_37_bench_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_37_bench_ref_n1.json 2> gpurun_out/r2_37_bench_ref_n1.err; echo ref rc=$?
PS_SPP=16 PS_REPS=3 PS_KERNEL=mega ncu --set full --import-source on --clock-control none -k regex:'k_mega_pixel' --launch-skip 2 --launch-count 1 -o /tmp/soup python tools/prof_soup.py > /dev/null 2>&1
(python tools/ncu_summary.py /tmp/soup.ncu-rep 2>/dev/null; echo; python tools/ncu_lines.py /tmp/soup.ncu-rep k_mega_pixelILi3ELb1ELi0ELb1E k_mega_pixel 40 2>&1) > gpurun_out/r2_38_grid_soup1m_mega_16spp_final.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_39_launches_bench_config4.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1
echo done
