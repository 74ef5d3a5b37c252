"""profiles/traffic.json from an .ncu-rep of the eager bench: dram__bytes_read.sum + dram__bytes_write.sum per launch of
every kernel (mean over the captured launches).  The tile kernel of the split pipeline is stored as k_raster_shade_split.
usage: python tools/ncu_traffic.py gpurun_out/prof_final.ncu-rep [split|fused] > profiles/traffic.json"""
import csv, json, re, subprocess, sys, collections
rep = sys.argv[1]
split = (sys.argv[2] if len(sys.argv) > 2 else "split") == "split"
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
def col(name): return hdr.index(name)
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
acc = collections.defaultdict(list)
for r in rows[2:]:
    name = re.sub(r'<.*', '', r[col('Kernel Name')].replace('void ', '').replace('<unnamed>::', ''))
    name = re.sub(r'\(.*', '', name)
    tot = 0.0
    for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        tot += float(r[col(m)]) * scale[units[col(m)]]
    acc[name].append(tot)
out = {}
for k, v in acc.items():
    out[k + ('_split' if split and k == 'k_raster_shade' else '')] = sum(v) / len(v)
json.dump(out, sys.stdout, indent=1)
print()
