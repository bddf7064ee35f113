"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    ms = v / 1e6 if unit in ('ns', 'nsecond') else (v / 1e3 if unit in ('us', 'usecond') else v)
    name = re.sub(r'\(.*', '', row['Kernel Name']); name = re.sub(r'<.*', '', name)[:80]
    agg[name][0] += 1; agg[name][1] += ms; tot += ms
print(f'total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches')
for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f'{ms:9.3f} ms {100*ms/tot:5.1f}% n={n:4d} {k}')
