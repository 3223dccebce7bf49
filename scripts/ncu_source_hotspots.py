"""Per-source-line hot spots of a kernel from an `ncu --set full --import-source on` capture:
    python scripts/ncu_source_hotspots.py <prof.ncu-rep> [top_n]
Aggregates `ncu --page source --print-source cuda,sass` by (file, line): warp-stall samples, executed warp
instructions, the dominant stall reasons."""
import collections
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    cur_file, hdr = None, None
    agg = collections.OrderedDict()
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or r[0] in ('Function Name',):
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr[2:], r[2:]))          # (the second 'Source' column is the SASS text or '-')
        def num(k):
            try:
                return float(d.get(k, '0') or 0)
            except ValueError:
                return 0.0
        if r[2] != '-':                        # SASS row: skip, the CUDA line row above already aggregates it
            continue
        key = (cur_file, line)
        a = agg.setdefault(key, {'src': r[1].strip()[:90], 'samples': 0.0, 'inst': 0.0, 'stalls': collections.Counter()})
        a['samples'] += num('# Samples')
        a['inst'] += num('Instructions Executed')
        for k in hdr:
            if k.startswith('stall_') and '(Not Issued)' not in k:
                a['stalls'][k] += num(k)
    tot_s = sum(a['samples'] for a in agg.values()) or 1.0
    tot_i = sum(a['inst'] for a in agg.values()) or 1.0
    print('total samples %.0f, total warp instructions %.3e' % (tot_s, tot_i))
    print('%-22s %6s %6s  %-38s %s' % ('file:line', 'smpl%', 'inst%', 'top stalls', 'source'))
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:top]:
        st = ' '.join('%s:%.0f%%' % (k[6:], 100 * v / max(a['samples'], 1)) for k, v in a['stalls'].most_common(3))
        print('%-22s %6.2f %6.2f  %-38s %s' % ('%s:%d' % (f, l), 100 * a['samples'] / tot_s, 100 * a['inst'] / tot_i, st, a['src']))


if __name__ == '__main__':
    main()
