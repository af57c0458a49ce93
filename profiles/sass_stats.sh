#!/bin/bash
# Static SASS statistics per kernel of an object / shared library: instructions, FP64 arithmetic, local-memory
# accesses, TMA / cp.async instructions.  Usage: profiles/sass_stats.sh <file.o|libsbk.so> [name filter]
f=$1; pat=${2:-.}
cuobjdump -sass "$f" | awk -v pat="$pat" '
  /Function :/ {fn=$3}
  /^[ \t]+\/\*[0-9a-f]+\*\/ / { n[fn]++;
      if ($0 ~ /DFMA|DMUL|DADD/) d[fn]++; if ($0 ~ /DFMA/) fma[fn]++;
      if ($0 ~ /LDL|STL/) l[fn]++; if ($0 ~ /UBLKCP/) t[fn]++; if ($0 ~ /LDGSTS/) g[fn]++;
      if ($0 ~ /MUFU/) m[fn]++; if ($0 ~ / LDG| STG/) gm[fn]++; if ($0 ~ / LDS| STS/) sm[fn]++ }
  END { printf "%8s %8s %8s %6s %6s %6s %6s %6s %6s  %s\n", "instr", "fp64", "dfma", "local", "ldg/stg", "lds/sts", "ublkcp", "ldgsts", "mufu", "kernel";
        for (k in n) if (k ~ pat) printf "%8d %8d %8d %6d %6d %6d %6d %6d %6d  %s\n", n[k], d[k], fma[k], l[k], gm[k], sm[k], t[k], g[k], m[k], k }'
