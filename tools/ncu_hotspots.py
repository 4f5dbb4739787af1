"""Top stall locations (SASS) of an .ncu-rep captured with --import-source on."""
import csv
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
body = rows[2:]
src, samp, ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(float(r[samp] or 0) for r in body)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print(f'total samples {tot:.0f}')
for idx, r in sorted(enumerate(body), key=lambda x: -float(x[1][samp] or 0))[:n]:
    top = sorted(((float(r[i] or 0), hdr[i]) for i in stalls), reverse=True)[:2]
    why = ' '.join(f'{h[6:]}={v:.0f}' for v, h in top if v > 0)
    print(f'{float(r[samp])/tot*100:6.2f}% line{idx:5d} exec={r[ex]:>10s}  {r[src][:90]:90s} {why}')
