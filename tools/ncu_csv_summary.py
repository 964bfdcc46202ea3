#!/usr/bin/env python
"""Summarise the raw-page CSV of an `ncu --set full` capture (tools/gpu_profile.sh) into profiles/:

    python tools/ncu_csv_summary.py <raw.csv> <out.json> <commit> [--traffic profiles/ncu_traffic.json]

One entry per captured launch with the metrics the bench line and DESIGN.md quote; with --traffic the per-kernel DRAM
bytes per launch (read + write) are merged into the table bench.py reads for `roofline.traffic`.
"""
import csv
import json
import statistics
import sys

KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
    'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
]
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'us': 1e-3, 'ns': 1e-6, 'ms': 1.0, 'msecond': 1.0, 'usecond': 1e-3, 'nsecond': 1e-6}


def classify(name, ms, grid):
    if 'conv_igemm_pair_fp4_kernel' in name:
        return 'conv_igemm_pair_fp4_kernel 3x3 512->512' if ms > 0.6 else 'conv_igemm_pair_fp4_kernel (1x1 or 256-channel layer)'
    for k in ('gn_apply_rows_kernel<0', 'gn_apply_rows_kernel<1', 'gn_apply_rows_kernel<2', 'dsac_score_kernel', 'dsac_refine_kernel',
              'dsac_sample_kernel'):
        if k in name:
            return k + ('>' if '<' in k else '')
    return name[:60]


def main():
    src, out, commit = sys.argv[1:4]
    rd = list(csv.reader(open(src)))
    head, units, rows = rd[0], rd[1], rd[2:]
    launches = []
    for r in rows:
        d = dict(zip(head, r))
        ent = {'kernel': d['Kernel Name'][:100]}
        for k in KEEP:
            if k in d and d[k] != '':
                v = float(d[k].replace(',', ''))
                u = units[head.index(k)]
                ent[k] = v * SCALE.get(u, 1.0) if ('bytes' in k or 'time' in k) else v
        ent['ms'] = ent.get('gpu__time_duration.sum')
        ent['dram_bytes'] = ent.get('dram__bytes_read.sum', 0) + ent.get('dram__bytes_write.sum', 0)
        ent['class'] = classify(ent['kernel'], ent['ms'], ent.get('launch__grid_size'))
        launches.append(ent)
    groups = {}
    for e in launches:
        groups.setdefault(e['class'], []).append(e)
    summary = {}
    for k, es in groups.items():
        summary[k] = {'launches': len(es), 'ms_median': statistics.median(e['ms'] for e in es),
                      'dram_bytes_per_launch_median': statistics.median(e['dram_bytes'] for e in es),
                      'dram_GBps': statistics.median(e['dram_bytes'] / (e['ms'] * 1e-3) / 1e9 for e in es),
                      'tensor_pipe_active_pct': statistics.median(e.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0) for e in es),
                      'fp64_pipe_active_pct': statistics.median(e.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 0) for e in es),
                      'l2_hit_pct': statistics.median(e.get('lts__t_sector_hit_rate.pct', 0) for e in es),
                      'registers': es[0].get('launch__registers_per_thread')}
    doc = {'source_csv': src, 'commit': commit, 'command': 'ncu --set full --clock-control none (tools/gpu_profile.sh, tools/profile_step.py: 32 frames 480x720, 256 hypotheses)',
           'note': 'per-launch values under ncu are cold-cache and serialised; times in ms, bytes in bytes', 'summary': summary, 'launches': launches}
    json.dump(doc, open(out, 'w'), indent=1)
    for k, v in summary.items():
        print('%-58s n=%2d  %.3f ms  %.1f MB  %.0f GB/s  tensor %.0f%%  fp64 %.0f%%  L2 hit %.0f%%' % (
            k, v['launches'], v['ms_median'], v['dram_bytes_per_launch_median'] / 1e6, v['dram_GBps'], v['tensor_pipe_active_pct'],
            v['fp64_pipe_active_pct'], v['l2_hit_pct']))
    if '--traffic' in sys.argv:
        path = sys.argv[sys.argv.index('--traffic') + 1]
        try:
            table = json.load(open(path))
        except FileNotFoundError:
            table = {}
        for k, v in summary.items():
            table[k] = {'dram_bytes_per_launch': v['dram_bytes_per_launch_median'], 'source': out, 'commit': commit, 'launches': v['launches']}
        json.dump(table, open(path, 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
