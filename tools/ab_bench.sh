#!/bin/bash
# A/B of two builds of libcrossloc_b200.so on ONE box: crossloc_b200/_C/ab/lib_<name>.so, alternating runs of the headline bench
orig=crossloc_b200/_C/libcrossloc_b200.so
cp $orig /tmp/lib_orig.so
for round in 1 2 3; do
  for v in "$@"; do
    cp crossloc_b200/_C/ab/lib_$v.so $orig
    python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); c=d['clocks']; km=d['kernel_ms']
print('$v', round(d['ms_per_step'],3), 'ms  clk', c['sm_mhz'], ' 3x3 %.4f  1x1 %.4f  conv2 %.3f conv3 %.3f conv4 %.3f' % (km['512/512/3/1']['avg_ms'], km['512/512/1/1']['avg_ms'], km['32/64/3/2']['ms_per_step'], km['64/128/3/2']['ms_per_step'], km['128/256/3/2']['ms_per_step']))"
  done
done
cp /tmp/lib_orig.so $orig
