#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum` launch list (csv log) per kernel."""
import collections, csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, iv, ig, ib = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size'), hdr.index('Block Size')
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ik].split('(lumcu::')[0].split('(')[0].replace('void ', '') if 'lumcu' in r[ik] else r[ik].split('(')[0].replace('void ', '')
    if 'pool_kernel' in r[ik] or 'wg_kernel' in r[ik]:
        name = r[ik].split('(lumcu::')[0].replace('void ', '')
    a = agg.setdefault((name, r[ig], r[ib]), [0, 0.0]); a[0] += 1; a[1] += float(r[iv])
tot = sum(v[1] for v in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 3 --no-cpu --no-demo")
print("# cold-cache, serialised launches: compare SHARES.  The capture covers setup + the value / kernel-time / instrumented passes")
print(f"# {len(rows) - 1} launches, {tot / 1e6:.3f} ms total")
print("kernel | grid | block | launches | total_ms | avg_us | share")
for (n, g, b), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n} | {g} | {b} | {c} | {t / 1e6:.3f} | {t / c / 1e3:.1f} | {100 * t / tot:.2f}%")
