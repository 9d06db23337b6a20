"""Per-kernel totals (time, DRAM bytes, tensor-pipe activity) from a long-format `ncu --metrics ... --csv` log of one train step."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    per.setdefault(r[0], {'name': r[4], 'grid': r[8]})[r[12]] = (float(r[14].replace(',', '')), r[13])
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
detail = len(sys.argv) > 2
for k, v in per.items():
    name = v['name'].replace('void ', '').replace('(anonymous namespace)::', '')
    key = name[:int(sys.argv[2])] if detail else name.split('(')[0][:70]
    t, tu = v.get('gpu__time_duration.sum', (0, 'ns'))
    t_ms = t / 1e6 if tu in ('ns', 'nsecond') else (t / 1e3 if tu in ('us', 'usecond') else t)
    b = 0.0
    for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        val, unit = v.get(m, (0, 'byte'))
        b += val * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    tp = v.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', (0, ''))[0]
    a = agg[key]; a[0] += 1; a[1] += t_ms; a[2] += b; a[3] += tp * t_ms
tot = sum(a[1] for a in agg.values())
print(f'total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches, DRAM {sum(a[2] for a in agg.values()) / 1e9:.2f} GB')
print(f'{"kernel":72s} {"n":>4s} {"ms":>8s} {"dramGB":>7s} {"GB/s":>6s} {"tensor%":>7s}')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(40)]:
    print(f'{k:72s} {a[0]:4d} {a[1]:8.3f} {a[2] / 1e9:7.2f} {a[2] / 1e6 / max(a[1], 1e-9):6.0f} {a[3] / max(a[1], 1e-9):7.1f}')
