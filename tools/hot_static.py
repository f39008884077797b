"""static SASS instructions per source region restricted to lines that are dynamically warm in an ncu report"""
import collections, csv, re, subprocess, sys, tempfile, os
rep = sys.argv[1]; pat = sys.argv[2]; thr_frac = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-5
out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.split('\n')))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']; hdr_idx.append(len(rows) + 2)
dyn = collections.Counter()
for s in range(len(hdr_idx) - 1):
    hi = hdr_idx[s]; hdr = rows[hi]; f = rows[hi - 2][1].split('/')[-1]; ie = hdr.index('Instructions Executed')
    for r in rows[hi + 1: hdr_idx[s + 1] - 2]:
        if len(r) <= ie or not r[0].isdigit(): continue
        try: dyn[(f, int(r[0]))] += int(r[ie]) if r[ie] else 0
        except ValueError: pass
T = sum(dyn.values())
d = tempfile.mkdtemp(); os.chdir(d)
subprocess.run("cuobjdump -xelf all /root/repo/manusim_b200/libmsim.so", shell=True, capture_output=True)
txt = subprocess.run("nvdisasm -g -c msim_kernel.sm_100a.cubin", shell=True, capture_output=True, text=True).stdout.split('\n')
cur_fn = None; cur = None; stat = collections.Counter()
for l in txt:
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if cur_fn and pat in cur_fn and re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): stat[cur] += 1
srcf = open('/root/repo/manusim_b200/csrc/msim_kernel.cu').read().split('\n')
marks = [(i + 1, s.strip()[:60]) for i, s in enumerate(srcf) if re.match(r'\s*(case OP_|__device__|template|__global__|default:)', s)]
reg = collections.Counter(); regdyn = collections.Counter()
for key, n in stat.items():
    if key is None: continue
    f, ln = key
    if dyn.get(key, 0) < T * thr_frac: continue
    name = f
    if f == 'msim_kernel.cu':
        name = 'top'
        for ml, ms in marks:
            if ml <= ln: name = f"{ml}:{ms}"
            else: break
    reg[name] += n; regdyn[name] += dyn[key]
print("warm static instrs:", sum(reg.values()), f"= {sum(reg.values())*16/1024:.1f} KB (threshold {thr_frac} of executed)")
for k, v in reg.most_common(40): print(f"{v:5d} static  {regdyn[k]/T*100:6.2f}% dyn  {k}")
