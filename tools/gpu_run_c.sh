#!/bin/bash
# round 2, run C: fused body-frame integrator: parity subset + throughput by register budget
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.txt 2>&1
tail -5 gpurun_out/r2c_pytest.txt
for mb in 2 3 4; do
  SBK_LOCAL_MINB=$mb SBK_TAG=minb$mb python tools/quick_perf.py pin_chain50_64k humanoid30_64k 2>&1 | tee -a gpurun_out/r2c_perf.txt
done
for wl in humanoid30_64k pin_chain50_64k; do
  SBK_SPL=4 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'tpiKernel' -s 1 -c 1 -o /tmp/prof_$wl python tools/quick_perf.py $wl > gpurun_out/r2c_prof_$wl.log 2>&1
  python profiles/summarize_ncu.py /tmp/prof_$wl.ncu-rep > gpurun_out/r2c_prof_$wl.txt
  python profiles/ncu_sass.py /tmp/prof_$wl.ncu-rep 30 > gpurun_out/r2c_prof_${wl}_sass.txt
  python profiles/ncu_lines.py /tmp/prof_$wl.ncu-rep 30 > gpurun_out/r2c_prof_${wl}_lines.txt
done
head -30 gpurun_out/r2c_prof_humanoid30_64k.txt
