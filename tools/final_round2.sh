# end-of-round measurements on one B200: full GPU test suite, our bench arm, ncu capture of the headline kernel (text summary only)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_41_bench_n1.json 2> gpurun_out/r2_41_bench_n1.err; echo bench rc=$?
PS_SPP=16 PS_REPS=3 PS_KERNEL=mega ncu --set full --import-source on --clock-control none -k regex:'k_mega_pixel' --launch-skip 2 --launch-count 1 -o /tmp/soup python tools/prof_soup.py > /dev/null 2>&1
(python tools/ncu_summary.py /tmp/soup.ncu-rep 2>/dev/null; echo; python tools/ncu_lines.py /tmp/soup.ncu-rep k_mega_pixelILi3ELb1ELi0ELb1E k_mega_pixel 40 2>&1) > gpurun_out/r2_42_grid_soup1m_mega_16spp_final.txt
echo done
