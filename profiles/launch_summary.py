"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
   python profiles/launch_summary.py gpurun_out/launches.csv [min_ms_for_real_jacobi]"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0, 0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '')
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v
        if name.startswith('k_jacobi_sweep') and v < 0.05:
            name = name + ' (skipped: converged)'
        a = agg[name]
        a[0] += 1
        a[1] += v
        tot += v
    print(f'total kernel time {tot:.2f} ms over {sum(a[0] for a in agg.values())} launches')
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f'{k:42s} n={a[0]:5d} total={a[1]:9.3f} ms  avg={a[1] / a[0]:8.4f} ms  share={100 * a[1] / tot:5.1f}%')


if __name__ == '__main__':
    main(sys.argv[1])
