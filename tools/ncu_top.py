"""Top stall sites of a kernel from an ncu report's source page (SASS view).  Usage: python tools/ncu_top.py report.ncu-rep [n]
Prints the instructions with the most warp-stall samples plus the stall-reason columns that are non-zero for them."""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
cs = hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') or h.startswith('Stall') or 'stall' in h.lower()]
total = sum(int(r[cs] or 0) for r in body)
print('total samples', total, 'instructions', len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][cs] or 0))[:n]
for i in sorted(order):
    r = body[i]
    reasons = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols if c != cs and (r[c] or '0').isdigit() and int(r[c]) > 0), reverse=True)[:4]
    print(f'{i:5d} {int(r[cs]):6d} {100.0 * int(r[cs]) / max(total, 1):5.1f}%  {r[1].strip()[:70]:70s} {reasons}')
