#!/bin/bash
# A/B of two builds of libcrossloc_b200.so on ONE box: crossloc_b200/_C/ab/lib_<name>.so, alternating runs of the headline bench
# (boxes differ by +-2 % in SM clock under the power cap, so only same-box comparisons resolve changes of a per cent)
orig=crossloc_b200/_C/libcrossloc_b200.so
cp $orig /tmp/lib_orig.so
for round in 1 2 3; do
  for v in "$@"; do
    cp crossloc_b200/_C/ab/lib_$v.so $orig
    python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 | python tools/ab_line.py $v
  done
done
cp /tmp/lib_orig.so $orig
