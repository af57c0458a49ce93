#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  big=40000; [ $tool != memcheck ] && big=38400
  SBK_SANITIZE_BIG=$big timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py > gpurun_out/r2_sanitizer_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_SUBSET_DONE|hazard|Invalid|Error" gpurun_out/r2_sanitizer_$tool.txt | head -12
done
