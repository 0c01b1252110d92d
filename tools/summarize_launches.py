"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X ...`) into a markdown table:
python tools/summarize_launches.py gpurun_out/launches.csv "<command line>" > profiles/rNN_launches_summary.md"""
import csv, sys
from collections import defaultdict
path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '')
rows = []
with open(path, newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = defaultdict(lambda: [0.0, 0])
for r in rd:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[iu], 1e-3)
    a = agg[r[ik]]
    a[0] += v; a[1] += 1
tot = sum(a[0] for a in agg.values()); n = sum(a[1] for a in agg.values())
ours = sum(a[0] for k, a in agg.items() if 'dv::' in k or k.startswith(('void dv', 'dv')) or 'slot_stream' in k or 'gemm_bf16' in k or 'flash_' in k or 'skinny_' in k or 'layernorm_' in k)
print(f'Command (B200, 1 GPU): `{cmd}`\n')
print(f'launches in window: {n}; summed device time {tot / 1e3:.1f} ms; share of hand-written kernels: {100 * ours / tot:.1f} %\n')
print('| share | time (us) | launches | kernel |\n|---|---|---|---|')
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f'| {100 * t / tot:.1f}% | {t:.0f} | {c} | `{k[:110]}` |')
