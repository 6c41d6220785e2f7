"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, im = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
iu = hdr.index('Metric Unit')
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[im] != 'gpu__time_duration.sum':
        continue
    v = float(r[iv].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(r[iu], 1.0)
    name = r[ik].split('(')[0]
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v for _, v in tot.values())
print(f'{"kernel":70s} {"launches":>8s} {"total us":>12s} {"avg us":>10s} {"share":>7s}')
for name, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{name[:70]:70s} {n:8d} {v:12.1f} {v / n:10.2f} {100 * v / total:6.2f}%')
print(f'{"all":70s} {sum(n for n, _ in tot.values()):8d} {total:12.1f}')
