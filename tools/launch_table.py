"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) by kernel name.
    python tools/launch_table.py gpurun_out/launches.csv [top]"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hdr = None
    agg = collections.OrderedDict()
    tot, n = 0.0, 0
    for r in rows:
        if hdr is None:
            if 'Kernel Name' in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        name = re.sub(r'\(.*', '', d['Kernel Name'])[:90]
        try:
            t = float(d['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        unit = d['Metric Unit']
        t = t / 1e3 if unit == 'ns' else t * 1e3 if unit == 'ms' else t
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
        n += 1
    print(f'{n} launches, {tot:.1f} us of kernel time')
    print(f'{"us":>9s} {"share":>6s} {"n":>5s} {"us each":>8s}  kernel')
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{t:9.1f} {100 * t / tot:5.1f}% {c:5d} {t / c:8.1f}  {name}')


if __name__ == '__main__':
    main()
