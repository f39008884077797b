"""static SASS instruction count per source line / line range of one kernel (I-cache budgeting)."""
import collections, re, subprocess, sys, tempfile, os
pat = sys.argv[1] if len(sys.argv) > 1 else 'ILi0ELi1'
d = tempfile.mkdtemp(); os.chdir(d)
subprocess.run("cuobjdump -xelf all /root/repo/manusim_b200/libmsim.so", shell=True, capture_output=True)
txt = subprocess.run("nvdisasm -g -c msim_kernel.sm_100a.cubin", shell=True, capture_output=True, text=True).stdout.split('\n')
cur_fn = None; cur_line = None
cnt = collections.defaultdict(collections.Counter)
for l in txt:
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur_line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l) and cur_fn: cnt[cur_fn][cur_line] += 1
fn = [f for f in cnt if pat in f][0]
c = cnt[fn]
print(fn, sum(c.values()), 'instr')
src = open('/root/repo/manusim_b200/csrc/msim_kernel.cu').read().split('\n')
# aggregate by 'case OP_' / function regions: find region start lines
marks = [(i + 1, s.strip()[:60]) for i, s in enumerate(src) if re.match(r'\s*(case OP_|__device__|template|__global__|default:)', s)]
reg = collections.Counter()
for (f, ln), v in c.items():
    if f != 'msim_kernel.cu': reg[f] += v; continue
    name = 'top'
    for ml, ms in marks:
        if ml <= ln: name = f"{ml}:{ms}"
        else: break
    reg[name] += v
for k, v in reg.most_common(70): print(f"{v:6d} {k}")
