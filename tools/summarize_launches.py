"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import csv
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        if unit in ('us', 'usecond'):
            v *= 1e3
        elif unit in ('ms', 'msecond'):
            v *= 1e6
        rows.append((r['Kernel Name'].split('(')[0], v, r.get('Grid Size', ''), r.get('Block Size', '')))
    tot = sum(r[1] for r in rows)
    agg = OrderedDict()
    for name, v, g, b in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    print('total %d launches, %.1f us (serialised, cold-cache per-launch times: compare shares, not absolutes)' % (len(rows), tot / 1e3))
    print('%-60s %6s %12s %7s' % ('kernel', 'count', 'total_us', 'share'))
    for name, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-60s %6d %12.1f %6.1f%%' % (name[:60], c, v / 1e3, 100 * v / tot))
    if '-v' in sys.argv:
        for i, (name, v, g, b) in enumerate(rows):
            print('%4d %-50s %10.1f us grid %s block %s' % (i, name[:50], v / 1e3, g, b))


if __name__ == '__main__':
    main(sys.argv[1])
