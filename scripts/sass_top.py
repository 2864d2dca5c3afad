#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv --print-source sass` export.
  python scripts/sass_top.py FILE.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
I = lambda r, k: int(float(r[ix[k]] or 0))
tot = sum(I(r, '# Samples') for r in data)
ex = sum(I(r, 'Instructions Executed') for r in data)
print('total samples', tot, 'sass lines', len(data), 'warp instructions', ex)
keys = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {k: sum(I(r, k) for r in data) for k in keys}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for pos, r in sorted(enumerate(data), key=lambda pr: -I(pr[1], '# Samples'))[:n]:
    st = {k: I(r, k) for k in keys}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(str(pos).rjust(5), str(I(r, '# Samples')).rjust(6), str(I(r, 'Instructions Executed')).rjust(8),
          r[ix['Source']][:80].ljust(80), top)
