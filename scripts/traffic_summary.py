"""Sum dram__bytes_{read,write} and durations per kernel from an ncu --csv metrics log -> JSON (per step)."""
import csv, collections, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1e-3, 'ms': 1, 'ns': 1e-6, 's': 1e3}
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0].replace('void ', '').replace('hebb::', '').split('<')[0]
    agg[name][r[mi]] += float(r[vi].replace(',', '')) * scale.get(r[ui], 1)
    if r[mi] == 'gpu__time_duration.sum':
        cnt[name] += 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
out = {k: dict(launches_per_step=cnt[k] / steps, dram_read_bytes_per_step=v['dram__bytes_read.sum'] / steps,
               dram_write_bytes_per_step=v['dram__bytes_write.sum'] / steps, ncu_ms_per_step=v['gpu__time_duration.sum'] / steps)
       for k, v in agg.items()}
json.dump(out, open(sys.argv[3], 'w'), indent=1, sort_keys=True) if len(sys.argv) > 3 else None
for k, v in out.items():
    print(k, {a: round(b, 3) if b < 1e6 else f'{b / 1e6:.1f}e6' for a, b in v.items()})
