"""Segments an ncu source-page CSV (one kernel) by execution count: python scripts/sass_hot.py file.csv"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hi = [k for k, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
isrc = hdr.index('Source'); isam = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
tot_s = sum(int(r[isam]) for r in data); tot_e = sum(int(r[iex]) for r in data)
print("total samples", tot_s, "total warp-inst", tot_e, "n sass", len(data))
segs = []; cur = None
for k, r in enumerate(data):
    e = int(r[iex]); s = int(r[isam])
    if cur and abs(e - cur['e']) <= 0.03 * max(e, cur['e'], 1):
        cur['n'] += 1; cur['s'] += s; cur['ex'] += e; cur['end'] = k
    else:
        cur = {'start': k, 'end': k, 'e': e, 'n': 1, 's': s, 'ex': e}; segs.append(cur)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
for sg in segs:
    if sg['ex'] > thr * tot_e or sg['s'] > thr * tot_s:
        ops = [data[k][isrc].split()[0] if not data[k][isrc].strip().startswith('@') else data[k][isrc].split()[1] for k in range(sg['start'], sg['end'] + 1)]
        c = Counter(ops).most_common(7)
        top = max(range(sg['start'], sg['end'] + 1), key=lambda k: int(data[k][isam]))
        print(f"sass[{sg['start']:5d}-{sg['end']:5d}] n={sg['n']:4d} exec/inst={sg['e']:8d} inst%={100*sg['ex']/tot_e:5.1f} samp%={100*sg['s']/tot_s:5.1f} {c}  hottest: {data[top][isrc].strip()[:50]} ({data[top][isam]})")
