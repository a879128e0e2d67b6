import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0,0.0])
for r in rows[hi+1:]:
    if len(r) <= vi: continue
    name = r[ki].split('(')[0][:70]
    v = float(r[vi].replace(',','')); unit = r[ui]
    if unit == 'ns': v /= 1e3
    elif unit == 'ms': v *= 1e3
    elif unit in ('s','second'): v *= 1e6
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print('kernel,launches,total_us,avg_us,share')
for k,v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%s,%d,%.1f,%.2f,%.4f' % (k, v[0], v[1], v[1]/v[0], v[1]/tot))
