"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py <csv> [divide-by]"""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
agg = collections.defaultdict(lambda: [0, 0.0]); seq = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = re.sub(r'\(.*', '', row['Kernel Name']); v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']; v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    agg[name][0] += 1; agg[name][1] += v; seq.append((name, v))
tot = sum(t for _, t in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:44s} n={n:6d} total={t / 1e3 / div:9.3f} ms  avg={t / n:9.1f} us  share={100 * t / tot:5.1f} %")
if len(sys.argv) > 3:
    xs = [v for n, v in seq if sys.argv[3] in n]
    print(len(xs), [round(x) for x in xs[:6]], [round(x) for x in xs[len(xs) // 2:len(xs) // 2 + 4]], [round(x) for x in xs[-4:]])
