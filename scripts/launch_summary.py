"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3}.get(row['Metric Unit'], v / 1e3)
    name = row['Kernel Name']
    name = re.sub(r'\(anonymous namespace\)::', '', name)
    name = re.sub(r'^void ', '', name)
    m = re.match(r'([A-Za-z0-9_:]+)(<[^(]*>)?', name)
    short = m.group(1) + ((m.group(2) or '')[:40]) if m else name[:70]
    tot[short] += v
    cnt[short] += 1
s = sum(tot.values())
print(f'total {s / 1e3:.3f} ms over {sum(cnt.values())} launches')
print(f'{"ms":>9} {"share":>6} {"n":>5}  kernel')
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f'{v / 1e3:9.3f} {100 * v / s:5.1f}% {cnt[k]:5d}  {k}')
