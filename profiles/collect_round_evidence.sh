#!/bin/bash
# round-1 evidence: bench lines, launch list, ncu captures (summaries only), trajectory agreement
python bench.py --steps 5 --warmup 3 --all-workloads > gpurun_out/r1_bench_all.json 2> gpurun_out/r1_bench_all.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_reference.json 2>> gpurun_out/r1_bench_all.err
for wl in pin_chain50_64k humanoid30_64k branched_tree1000_256; do python bench.py --impl reference --workload $wl --steps 1 --warmup 1 >> gpurun_out/r1_bench_reference.json 2>> gpurun_out/r1_bench_all.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-workloads > gpurun_out/launch_bench.log 2>&1
for wl in double_pendulum_1M pin_chain50_64k humanoid30_64k branched_tree1000_256; do
  spl=4; [ $wl = double_pendulum_1M ] && spl=20; [ $wl = branched_tree1000_256 ] && spl=1
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'tpiKernel<\(int\)7|fusedRkmKernel|glRkmKernel' -s 2 -c 1 -o /tmp/prof_$wl python bench.py --workload $wl --steps 1 --warmup 3 --steps-per-launch $spl --no-cpu-baseline --no-extra-workloads > gpurun_out/prof_$wl.log 2>&1
  python profiles/summarize_ncu.py /tmp/prof_$wl.ncu-rep > gpurun_out/r1_prof_$wl.txt
  python profiles/ncu_sass.py /tmp/prof_$wl.ncu-rep 25 > gpurun_out/r1_prof_${wl}_sass.txt
  python profiles/ncu_lines.py /tmp/prof_$wl.ncu-rep 25 > gpurun_out/r1_prof_${wl}_lines.txt
done
python profiles/trajectory_agreement.py > gpurun_out/r1_trajectory_agreement.json 2> gpurun_out/traj.err
ls -la gpurun_out
