"""Top stall sites of one launch in an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
hdr = rows[hi]
si, col = hdr.index('Source'), hdr.index('# Samples')
data, seen = [], set()
for r in rows[hi + 1:]:
    try:
        key = r[0]
        if key in seen:
            continue
        seen.add(key)
        data.append((int(r[col]), r[si]))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
print('total samples', tot)
for n, s in sorted(data, key=lambda x: -x[0])[:top]:
    print(f'{n:7d} {100 * n / tot:5.1f}%  {s[:120]}')
