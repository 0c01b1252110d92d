"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list: the LAST
training step (from its first patch-embedding launch to the end of the list), grouped by kernel family: launches, summed
device time, share, DRAM bytes.  Optionally writes the GEMM family's DRAM traffic per launch as JSON (bench.py roofline.traffic).
Usage: python tools/launch_summary.py launches.csv [--traffic-json out.json --command "<the ncu command>"]"""
import collections, csv, json, re, sys
args = sys.argv[1:]
path = args[0]
tj = args[args.index('--traffic-json') + 1] if '--traffic-json' in args else None
cmd = args[args.index('--command') + 1] if '--command' in args else ''
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]
kn, mn, mv, un = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
launch = collections.OrderedDict()     # id -> {name, us, rd, wr}
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[0].isdigit():
        continue
    d = launch.setdefault(int(r[0]), {'name': r[kn], 'us': 0.0, 'rd': 0.0, 'wr': 0.0})
    v = float(r[mv].replace(',', ''))
    if r[mn] == 'gpu__time_duration.sum':
        d['us'] = v * {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3}.get(r[un], 1e-3)
    elif r[mn].startswith('dram__bytes_'):
        f = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r[un], 1)
        d['rd' if 'read' in r[mn] else 'wr'] = v * f
L = list(launch.values())
pos = [i for i, d in enumerate(L) if 'patch_embed_fwd' in d['name'] or 'patchify' in d['name']]
# the last TRAINING step starts at the last forward patch embedding that is followed by a backward (an AdamW launch after it)
starts = [i for i in pos if 'patch_embed_fwd' in L[i]['name']] or pos
lo = starts[-1]
step = L[lo:]
def fam(n):
    n = re.sub(r'<.*', '', n); n = re.sub(r'\(.*', '', n).replace('void ', '')
    if n.startswith('dv::'): return n
    if 'nccl' in n.lower(): return 'nccl'
    if 'at::' in n or 'at_cuda' in n or 'elementwise' in n or 'reduce_kernel' in n: return 'torch: ' + n.split('::')[-1][:44]
    return 'other: ' + n[:50]
tot = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in step:
    t = tot[fam(d['name'])]; t[0] += 1; t[1] += d['us']; t[2] += d['rd']; t[3] += d['wr']
T = sum(t[1] for t in tot.values())
dv = [(k, t) for k, t in tot.items() if k.startswith('dv::')]
print(f'step: {len(step)} launches, summed device time {T / 1e3:.3f} ms (serialised, cold-cache ncu timing: compare SHARES)')
print(f'hand-written dv:: kernels: {sum(t[0] for _, t in dv)} launches, {sum(t[1] for _, t in dv) / 1e3:.3f} ms ({100 * sum(t[1] for _, t in dv) / T:.1f} %); '
      f'torch / library: {len(step) - sum(t[0] for _, t in dv)} launches, {(T - sum(t[1] for _, t in dv)) / 1e3:.3f} ms '
      f'({100 * (T - sum(t[1] for _, t in dv)) / T:.1f} %)')
print('\n| kernel family | launches | time (us) | share | DRAM read (MB) | DRAM write (MB) |\n|---|---|---|---|---|---|')
for k, t in sorted(tot.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f'| `{k}` | {t[0]} | {t[1]:.1f} | {100 * t[1] / T:.1f} % | {t[2] / 1e6:.1f} | {t[3] / 1e6:.1f} |')
if tj:
    g = tot['dv::gemm_bf16_kernel']
    json.dump({'dram_bytes_per_launch': (g[2] + g[3]) / max(g[0], 1), 'launches_captured': g[0],
               'dram_read_bytes_total': g[2], 'dram_write_bytes_total': g[3],
               'source': f'{path} -- every gemm_bf16_kernel launch (forward, dgrad, wgrad) of one training step; command: {cmd}'},
              open(tj, 'w'), indent=1)
