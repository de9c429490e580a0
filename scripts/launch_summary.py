"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    if unit.startswith('n'): v /= 1e3
    elif unit.startswith('m'): v *= 1e3
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('<unnamed>::', '')
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f'# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised: compare SHARES)')
print(f'# {"total_us":>10} {"share":>6} {"n":>4} {"avg_us":>8}  kernel')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'  {t:10.1f} {100 * t / tot:5.1f}% {n:4d} {t / n:8.1f}  {k[:100]}')
