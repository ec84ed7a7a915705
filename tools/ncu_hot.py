"""Per-instruction stall samples of a kernel's hot loop from an .ncu-rep captured with --import-source on.
usage: python tools/ncu_hot.py report.ncu-rep [min_exec_fraction] [min_samples]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.03
mins = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
I = lambda r, k: int(r[ix[k]] or 0)
tot = sum(I(r, '# Samples') for r in data)
mx = max(I(r, 'Instructions Executed') for r in data)
print(rows[0][1][:90]); print('total samples', tot, 'max exec', mx)
cols = [h for h in hdr if h.startswith('stall_') and '(' not in h]
agg = {}
for r in data:
    ex, s = I(r, 'Instructions Executed'), I(r, '# Samples')
    for c in cols: agg[c] = agg.get(c, 0) + I(r, c)
    if ex > mx * frac and s >= mins:
        print(r[ix['Address']][-5:], '%9d' % ex, '%6d' % s, ' '.join('%s=%s' % (c[6:10], r[ix[c]]) for c in cols if r[ix[c]] not in ('0', '')), '|', r[ix['Source']][:72])
print({k: v for k, v in agg.items() if v})
