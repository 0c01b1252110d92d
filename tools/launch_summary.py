"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the LAST training step (from its patchify launch to the
end of the list or the next patchify), grouped by kernel family.  Usage: python tools/launch_summary.py launches.csv [step_index]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].isdigit()]
kn, mv, un = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
names = [r[kn] for r in body]
def us(r):
    v = float(r[mv].replace(',', ''))
    return v / 1e3 if r[un] in ('ns', 'nsecond') else v if r[un] in ('us', 'usecond') else v * 1e3
vals = [us(r) for r in body]
pos = [i for i, n in enumerate(names) if 'patchify' in n]
k = int(sys.argv[2]) if len(sys.argv) > 2 else len(pos) - 1
lo, hi2 = pos[k], (pos[k + 1] if k + 1 < len(pos) else len(names))
def fam(n):
    n = re.sub(r'<.*', '', n); n = re.sub(r'\(.*', '', n)
    n = n.replace('void ', '')
    if n.startswith('dv::'): return n
    if 'at::native' in n or 'at_cuda' in n or 'elementwise' in n or 'reduce_kernel' in n: return 'torch: ' + n.split('::')[-1][:40]
    return 'other: ' + n[:50]
tot = collections.Counter(); cnt = collections.Counter()
for n, v in zip(names[lo:hi2], vals[lo:hi2]):
    f = fam(n); tot[f] += v; cnt[f] += 1
T = sum(tot.values())
print(f'step launches {hi2 - lo}, summed device time {T / 1e3:.3f} ms')
dvt = sum(v for f, v in tot.items() if f.startswith('dv::')); dvc = sum(c for f, c in cnt.items() if f.startswith('dv::'))
print(f'dv:: kernels {dvc} launches {dvt / 1e3:.3f} ms ({100 * dvt / T:.1f} %), torch/library {hi2 - lo - dvc} launches {(T - dvt) / 1e3:.3f} ms ({100 * (T - dvt) / T:.1f} %)')
for f, v in tot.most_common(40):
    print(f'{v:10.1f} us {100 * v / T:5.1f} %  x{cnt[f]:4d}  {f}')
