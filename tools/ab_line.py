"""One line of an A/B run: reads the bench JSON line on stdin (tools/ab_bench.sh)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
c, km = d['clocks'], d['kernel_ms']
gn = sum(v['ms_per_step'] for k, v in km.items() if k.startswith('gn_apply'))
print(sys.argv[1], '%.3f ms  clk %s  3x3 %.4f  1x1 %.4f  gn512 %.4f  gn512+res %.4f  gn_all %.3f  conv2 %.3f  e2e %.1f' % (
    d['ms_per_step'], c['sm_mhz'], km['512/512/3/1']['avg_ms'], km['512/512/1/1']['avg_ms'], km['gn_apply/512/1/0']['avg_ms'],
    km['gn_apply/512/1/1']['avg_ms'], gn, km['32/64/3/2']['ms_per_step'], d['e2e']['value']))
