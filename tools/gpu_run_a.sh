#!/bin/bash
# round 2, run A: parity suite on the body-frame integrator, old-vs-new throughput, ncu of the new kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.txt 2>&1
tail -5 gpurun_out/r2a_pytest.txt
python tools/quick_perf.py pin_chain50_64k humanoid30_64k > gpurun_out/r2a_perf_local.txt 2>&1
SBK_NOLOCAL=1 python tools/quick_perf.py pin_chain50_64k humanoid30_64k > gpurun_out/r2a_perf_old.txt 2>&1
cat gpurun_out/r2a_perf_local.txt gpurun_out/r2a_perf_old.txt
for wl in humanoid30_64k pin_chain50_64k; do
  SBK_SPL=4 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'tpiKernel' -s 1 -c 1 -o /tmp/prof_$wl python tools/quick_perf.py $wl > gpurun_out/r2a_prof_$wl.log 2>&1
  python profiles/summarize_ncu.py /tmp/prof_$wl.ncu-rep > gpurun_out/r2a_prof_$wl.txt
  python profiles/ncu_sass.py /tmp/prof_$wl.ncu-rep 30 > gpurun_out/r2a_prof_${wl}_sass.txt
  python profiles/ncu_lines.py /tmp/prof_$wl.ncu-rep 30 > gpurun_out/r2a_prof_${wl}_lines.txt
done
head -40 gpurun_out/r2a_prof_humanoid30_64k.txt
