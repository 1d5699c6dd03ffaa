# ncu captures for the issue-slot roofline of configs 3 and 5 (GPU box); only the text summaries are kept
QC_SPP=32 ncu --set full --clock-control none -k regex:'k_mega_pixel' --launch-skip 3 --launch-count 1 -o /tmp/a python tools/quick_c3.py > /dev/null 2>&1
QC_SPP=32 ncu --set full --clock-control none -k regex:'k_mega_pixel' --launch-skip 7 --launch-count 1 -o /tmp/b python tools/quick_c3.py > /dev/null 2>&1
PS_W=3840 PS_H=2160 PS_SPP=16 PS_REPS=3 PS_KERNEL=mega ncu --set full --clock-control none -k regex:'k_mega_pixel' --launch-skip 2 --launch-count 1 -o /tmp/c python tools/prof_soup.py > /dev/null 2>&1
python tools/ncu_summary.py /tmp/a.ncu-rep > gpurun_out/r2_34_base_1920x1080_32spp.txt 2>/dev/null
python tools/ncu_summary.py /tmp/b.ncu-rep > gpurun_out/r2_35_torus_1920x1080_32spp.txt 2>/dev/null
python tools/ncu_summary.py /tmp/c.ncu-rep > gpurun_out/r2_36_gridsoup1m_3840x2160_16spp.txt 2>/dev/null
echo done
