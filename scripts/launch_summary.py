"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = {'us': v / 1e3, 'ns': v / 1e6, 'ms': v, 's': v * 1e3}.get(r[ui], v)
    name = r[ki][:90]
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f'total {T:.2f} ms over {sum(cnt.values())} launches')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f'{v:9.2f} ms {100 * v / T:5.1f}% n={cnt[k]:5d}  {k}')
