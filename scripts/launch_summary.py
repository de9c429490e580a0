"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
usage: launch_summary.py launches.csv [--last-step]   (--last-step: only the launches of the final train step, which starts at
the last add_i64_kernel launch = the BN num_batches_tracked bump that opens TSRNTrainStep._fwd_bwd)"""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
def short(r):
    n = r['Kernel Name'].replace('(anonymous namespace)::', '').replace('<unnamed>::', '').replace('void ', '')
    m = re.match(r'([\w:]+(<[^(]*>)?)', n)
    return (m.group(1) if m else n)[:90]
if '--last-step' in sys.argv:
    starts = [i for i, r in enumerate(rows) if 'add_i64' in r['Kernel Name']]
    rows = rows[starts[-1]:]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in rows:
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    if unit.startswith('n'): v /= 1e3
    elif unit.startswith('m'): v *= 1e3
    name = short(row)
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f'# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised: compare SHARES)')
print(f'# {"total_us":>10} {"share":>6} {"n":>4} {"avg_us":>8}  kernel')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'  {t:10.1f} {100 * t / tot:5.1f}% {n:4d} {t / n:8.1f}  {k}')
