"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals of ONE
step (the launches between the last two `point_sampling_kernel`s).  python tools/launch_list.py file.csv"""
import collections
import csv
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(row['Metric Unit'], 1e-3)
        rows.append((row['Kernel Name'], v))
    marks = [i for i, (k, _) in enumerate(rows) if 'point_sampling' in k]
    step = rows[marks[-2]:marks[-1]] if len(marks) >= 2 else rows
    agg = collections.OrderedDict()
    for k, v in step:
        a = agg.setdefault(k[:100], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in step)
    ours = sum(v for k, v in step if 'msda::' in k)
    print(f'{len(rows)} launches captured; one step = {len(step)} launches, {tot:.1f} us of kernel time, '
          f'{100 * ours / max(tot, 1e-9):.1f} % in msda:: kernels')
    print('| kernel | launches | us total | share |\n|---|---:|---:|---:|')
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'| `{k}` | {n} | {v:.1f} | {100 * v / tot:.1f}% |')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
