#!/bin/bash
# Static SASS instruction counts per source line for one kernel: profiles/sass_lines.sh <file.o> <kernel-name regex> [top N]
f=$1; pat=$2; top=${3:-40}
d=$(mktemp -d); (cd $d && cuobjdump -xelf all "$(readlink -f $f)" > /dev/null && nvdisasm -g -c *.cubin 2>/dev/null) | awk -v pat="$pat" '
  /^\/\/--------------------- \.text\./ {fn=$2}
  fn ~ pat { if ($0 ~ /\/\/## File/) { loc=$0; sub(/.*\//, "", loc); sub(/", line /, ":", loc); sub(/ .*/, "", loc) }
             else if ($0 ~ /^[ \t]+\/\*[0-9a-f]+\*\//) { n[loc]++; if ($0 ~ /DFMA|DMUL|DADD/) dd[loc]++ } }
  END { for (k in n) print n[k], dd[k]+0, k }' | sort -rn | head -$top
rm -rf $d
