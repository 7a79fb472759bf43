"""Stall-sample summary of one kernel from `ncu -i X.ncu-rep --page source --csv` output (SASS view)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print(sorted(agg.items(), key=lambda x: -x[1])[:10])
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for i in top:
    r = data[i]
    st = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(i, r[ix['# Samples']], r[ix['Source']][:70], st)
