"""Print the last forward pass of an ncu launch-list CSV (per-kernel cold-cache times)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == 'ID':
        h = i
        break
rs = [r for r in rows[h + 1:] if len(r) >= 15]
idx = [i for i, r in enumerate(rs) if 'preprocess' in r[4]]
last = rs[idx[-1]:] if idx else rs
tot = 0
for r in last:
    print('%-28s grid %-16s %8.2f us' % (r[4].split('(')[0][:28], r[8], float(r[-1]) / 1e3))
    tot += float(r[-1]) / 1e3
print('total', tot)
