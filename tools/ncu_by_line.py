"""Join the ncu SASS source page with nvdisasm line info -> instructions / stall samples per CUDA source line.
usage: ncu_by_line.py report.ncu-rep <kernel-substring e.g. k_raster_shadeILi3> [top]"""
import csv, subprocess, sys, re, collections, os, tempfile, glob
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'latent-nerf-test_b200', 'liblp_b200.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
dis = subprocess.run(['nvdisasm', '-g', '-c', glob.glob(tmp + '/*.cubin')[0]], capture_output=True, text=True).stdout.splitlines()
lines_of = []   # per instruction (in order) -> source line
inside = False; cur = None
for ln in dis:
    if ln.startswith('//-----') and '.text.' in ln:
        inside = kern in ln; cur = None; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith('lp_b200.cu') else cur; continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s', ln):
        lines_of.append(cur)
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + re.sub(r'ILi\d+.*', '', kern)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; per = collections.defaultdict(lambda: [0, 0]); idx = 0; ninst = 0
first = True
for r in rows:
    if 'Source' in r and 'Instructions Executed' in r:
        if hdr is not None: break      # only the first captured launch
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    try: n = int(d['Instructions Executed']); st = int(d['Warp Stall Sampling (All Samples)'])
    except ValueError: continue
    line = lines_of[idx] if idx < len(lines_of) else None
    idx += 1
    per[line][0] += n; per[line][1] += st
src = open(os.path.join(os.path.dirname(so), 'csrc', 'lp_b200.cu')).read().splitlines()
tot = sum(v[0] for v in per.values()); ts = sum(v[1] for v in per.values())
print(f'{kern}: {idx} SASS instructions joined ({len(lines_of)} in cubin); total executed {tot}, stall samples {ts}')
for line, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[line - 1].strip()[:110] if line else '?'
    print(f'{v[0]/tot*100:5.1f}% inst {v[1]/max(ts,1)*100:5.1f}% stall | L{line} {text}')
