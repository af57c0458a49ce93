#!/bin/bash
mkdir -p gpurun_out
wl=branched_tree1000_256
SBK_CLUSTER=8 SBK_SPL=2 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'ctreeRkmKernel' -s 1 -c 1 -o /tmp/prof_$wl python tools/quick_perf.py $wl > gpurun_out/r2h_prof_$wl.log 2>&1
python profiles/summarize_ncu.py /tmp/prof_$wl.ncu-rep > gpurun_out/r2h_prof_$wl.txt
python profiles/ncu_sass.py /tmp/prof_$wl.ncu-rep 40 > gpurun_out/r2h_prof_${wl}_sass.txt
python profiles/ncu_lines.py /tmp/prof_$wl.ncu-rep 40 > gpurun_out/r2h_prof_${wl}_lines.txt
head -32 gpurun_out/r2h_prof_$wl.txt
