"""Summarise ncu outputs brought back in gpurun_out/ into small JSON files under profiles/.

  python scripts/ncu_summary.py launches <launches.csv> <out.json>     # per-kernel time shares of one step
  python scripts/ncu_summary.py full <prof.ncu-rep> <out.json> [edges] # key counters of a --set full capture
"""
import csv
import collections
import io
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.per_cycle_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_srcunit_tex_op_red.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split('(')[0]
        ms = float(r[vi].replace(',', '')) / {'ns': 1e6, 'us': 1e3, 'ms': 1.0}[r[ui]]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(v for _, v in agg.values())
    res = {'note': 'ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare '
                   'shares, not absolutes)', 'launches': len(rows) - 1, 'total_ms': tot,
           'kernels': [{'kernel': k, 'launches': c, 'ms': round(v, 4), 'share': round(v / tot, 4)}
                       for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])]}
    json.dump(res, open(out, 'w'), indent=1)
    for k in res['kernels'][:12]:
        print('%-60s n=%3d %8.3f ms %5.1f%%' % (k['kernel'][:60], k['launches'], k['ms'], 100 * k['share']))


def full(path, out, edges=None):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {'Kernel Name': r[hdr.index('Kernel Name')]}
        for h, u, v in zip(hdr, units, r):
            if h in KEYS:
                d['%s [%s]' % (h, u)] = v
        if edges:
            rd = float(d.get('dram__bytes_read.sum [Gbyte]', 0)) * 1e9
            wr = float(d.get('dram__bytes_write.sum [Gbyte]', 0)) * 1e9
            d['edges_per_launch'] = edges
            d['dram_bytes_per_edge'] = round((rd + wr) / edges, 2)
        res.append(d)
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps(res, indent=1)[:1500])


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else None)
