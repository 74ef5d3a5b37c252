"""Aggregate the ncu source page by CUDA source line (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, collections
path, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', f'regex:{kern}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = collections.OrderedDict()
cur = None
for r in rows:
    if 'Source' in r and 'Instructions Executed' in r:
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    try:
        n = int(d['Instructions Executed']); st = int(d['Warp Stall Sampling (All Samples)'])
    except ValueError:
        continue
    key = d.get('Source', '')
    a = agg.setdefault(key, [0, 0]); a[0] += n; a[1] += st
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print('total inst', tot, 'stall samples', tots, 'distinct', len(agg))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{a[0]/tot*100:5.1f}% inst {a[1]/max(tots,1)*100:5.1f}% stall | {k[:150]}')
