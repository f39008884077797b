"""instructions per SimPy event by source region, from an ncu report with SourceCounters (+ -lineinfo)"""
import csv, collections, re, subprocess, sys, json
rep, events = sys.argv[1], float(sys.argv[2])
out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.split('\n')))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
hdr_idx.append(len(rows) + 2)
tot = collections.Counter(); src = {}
for s in range(len(hdr_idx) - 1):
    hi = hdr_idx[s]; hdr = rows[hi]
    f = rows[hi - 2][1].split('/')[-1]
    ie = hdr.index('Instructions Executed')
    for r in rows[hi + 1: hdr_idx[s + 1] - 2]:
        if len(r) <= ie or not r[0].isdigit(): continue
        try: v = int(r[ie]) if r[ie] else 0
        except ValueError: v = 0
        tot[(f, int(r[0]))] += v; src[(f, int(r[0]))] = r[1]
T = sum(tot.values())
print("instr/event", T / events)
srcf = open('/root/repo/manusim_b200/csrc/msim_kernel.cu').read().split('\n')
marks = [(i + 1, s.strip()[:70]) for i, s in enumerate(srcf) if re.match(r'\s*(case OP_|__device__|template|__global__|default:)', s)]
reg = collections.Counter()
for (f, ln), v in tot.items():
    if f != 'msim_kernel.cu': reg[f] += v; continue
    name = 'top'
    for ml, ms in marks:
        if ml <= ln: name = f"{ml}:{ms}"
        else: break
    reg[name] += v
for k, v in reg.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30): print(f"{v/events:7.2f}  {k}")
if len(sys.argv) > 4:
    lo, hi = map(int, sys.argv[4].split('-'))
    for (f, ln), v in sorted(tot.items()):
        if f == 'msim_kernel.cu' and lo <= ln <= hi and v: print(f"{ln:5d} {v/events:6.2f}  {src[(f, ln)].strip()[:100]}")
