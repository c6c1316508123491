"""print the hot SASS instructions (stall samples) of an `ncu --page source --csv` dump: python scripts/ncu_hot.py file.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print("total samples", tot)
keys = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for n, r in enumerate(data):
    s = int(r[ix['# Samples']] or 0)
    if s >= tot * thr / 100:
        st = sorted(((int(r[ix[k]] or 0), k) for k in keys), reverse=True)[:3]
        print(n, r[ix['Source']][:72].ljust(72), s, "%.1f%%" % (100 * s / tot), r[ix['Instructions Executed']], ' '.join(f"{k[6:]}={v}" for v, k in st if v))
