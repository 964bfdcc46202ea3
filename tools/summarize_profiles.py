#!/usr/bin/env python
"""Turn raw ncu captures (gpurun_out/) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py launches <launches.csv> <warmup+steps> <out.md> [title] [marker kernel]
    python tools/summarize_profiles.py full <report.ncu-rep> <out.json> <kernel description> <command>

`launches`: per-kernel totals of an `ncu --metrics gpu__time_duration.sum` launch list; the capture covers the
allocation / packing launches of the first step too, so shares are over the whole capture and the per-step
column divides by the number of steps the command ran.
`full`: selected metrics of one `ncu --set full` kernel capture (read with `ncu -i ... --page raw --csv`).
"""
import csv
import io
import json
import re
import subprocess
import sys

FULL_METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'smsp__cycles_active.avg',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
    'lts__t_bytes.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__sass_thread_inst_executed_op_ffma_pred_on.sum', 'sm__inst_executed.sum', 'smsp__inst_executed.sum',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'sm__maximum_warps_per_active_cycle_pct',
    'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
]


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    name = name.replace('cl::', '').replace('(anonymous namespace)::', '')
    return name[:60]


def launches(path, steps, out, title, marker=None):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(io.StringIO(''.join(lines))):
        if r['Metric Name'] == 'gpu__time_duration.sum':
            rows.append((short(r['Kernel Name']), float(r['Metric Value']) * (1e-6 if r['Metric Unit'] == 'ns' else 1.0)))
    if marker:
        # steady state only: the launches between the first and the last launch of `marker` (one per step), i.e. whole
        # steps without the allocation / filter-packing launches of the first one
        idx = [i for i, (k, _) in enumerate(rows) if marker in k]
        if len(idx) >= 2:
            rows = rows[idx[0] + 1:idx[-1] + 1]
            steps = len(idx) - 1
    agg = {}
    for k, ms in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(v[1] for v in agg.values())
    with open(out, 'w') as fh:
        fh.write('# %s\n\n' % title)
        fh.write('Raw list: `%s` (%d launches captured over %d steps incl. warm-up; per-launch times under ncu are '
                 'serialised and cold-cache: compare SHARES with the live CUDA-event numbers of the bench line).\n\n'
                 % (path, len(rows), steps))
        fh.write('| kernel | launches | launches / step | ms total | ms / step | share |\n|---|---:|---:|---:|---:|---:|\n')
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write('| `%s` | %d | %.1f | %.3f | %.3f | %.1f%% |\n' % (k, n, n / steps, ms, ms / steps, 100 * ms / total))
        fh.write('| **total** | %d | %.1f | %.3f | %.3f | 100%% |\n' % (len(rows), len(rows) / steps, total, total / steps))


def full(report, out, kernel, command):
    txt = subprocess.run(['ncu', '-i', report, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    header, units, values = rd[0], rd[1], rd[2]
    metrics = {}
    for name in FULL_METRICS:
        if name in header:
            i = header.index(name)
            metrics[name] = {'value': values[i], 'unit': units[i]}

    def num(name):
        m = metrics.get(name)
        if not m:
            return None
        v = float(m['value'].replace(',', ''))
        scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}.get(m['unit'], 1.0)
        return v * scale
    rd_b, wr_b = num('dram__bytes_read.sum'), num('dram__bytes_write.sum')
    doc = {'kernel': kernel, 'command': command, 'report': report,
           'dram_bytes_per_launch': (rd_b + wr_b) if rd_b is not None and wr_b is not None else None,
           'metrics': metrics}
    with open(out, 'w') as fh:
        json.dump(doc, fh, indent=1)
    print(json.dumps({k: v['value'] + ' ' + v['unit'] for k, v in metrics.items()}, indent=1))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else 'ncu launch list',
                 sys.argv[6] if len(sys.argv) > 6 else None)
    elif sys.argv[1] == 'full':
        full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
    else:
        raise SystemExit(__doc__)
