#!/usr/bin/env python
"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def main(path, top=16):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = d['Kernel Name'][:64]
        v = float(d['Metric Value'].replace(',', ''))
        u = d['Metric Unit']
        v = v / 1e3 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1e3)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print('%-64s n=%3d %10.1f us %5.1f%%' % (k, n, t, 100 * t / tot))
    print('total %.1f us over %d kernels' % (tot, sum(a[0] for a in agg.values())))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16)
