"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` log into per-kernel totals (count, total us, mean us), in
launch order of first appearance.  usage: python scripts/ncu_launches.py launches.csv [skip_first_n]"""
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        rows.append((int(r['ID']), name, us))
    for i, n, us in rows:
        print(f'{i:5d} {us:10.1f} us  {n}')


if __name__ == '__main__':
    main()
