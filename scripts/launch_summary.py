"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: last N launches, per kernel."""
import csv, sys
path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 34
rows = list(csv.DictReader([l for l in open(path) if not l.startswith('==')]))
tot = 0.0
for r in rows[-n:]:
    us = float(r['Metric Value']) / 1e3
    tot += us
    name = r['Kernel Name'].split('(')[0].replace('void ', '').replace('unnamed>::', '')
    print(f"{us:9.1f} us  {name:40s} grid {r['Grid Size']:18s} block {r['Block Size']}")
print(f"{tot:9.1f} us  TOTAL of last {n}")
