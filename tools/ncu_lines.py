"""
Development tool: per-source-line instruction / stall-sample shares of one kernel from an ncu report.
  python tools/ncu_lines.py report.ncu-rep object.o kernel_mangled_substring [top]
Joins `ncu --page source --csv` (SASS order, with counters) with `nvdisasm -g` (SASS order, with line info) of the same object.
"""
import csv, io, os, re, subprocess, sys, tempfile

rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, inside = [], None, False
for l in dis:
	if l.startswith('//---') and '.text.' in l:
		inside = kern in l
		continue
	if not inside:
		continue
	m = re.search(r'//## File "([^"]+)", line (\d+)', l)
	if m:
		cur = (os.path.basename(m.group(1)), int(m.group(2)))
		continue
	if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
		lines.append((cur, l.split('*/', 1)[1].strip()))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the report may hold several kernels: take the block whose name line holds the (demangled) kernel, else the first
start = 0
for i, r in enumerate(rows):
	if r and r[0] == 'Address':
		start = i
		break
hdr = rows[start]
iS, iN, iI, iT = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
sass = []
for r in rows[start + 1:]:
	if len(r) < len(hdr) or r[0] == 'Address':
		break
	sass.append((r[iS].strip(), int(r[iN] or 0), int(r[iI] or 0), int(r[iT] or 0)))
if len(sass) != len(lines):
	print('warning: instruction counts differ', len(sass), len(lines), file=sys.stderr)
agg = {}
for (loc, txt), (s, smp, inst, tinst) in zip(lines, sass):
	a = agg.setdefault(loc, [0, 0, 0, 0])
	a[0] += smp; a[1] += inst; a[2] += tinst; a[3] += 1
tot_s, tot_i = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f'total warp instructions {tot_i}, stall samples {tot_s}, sass instructions {len(sass)}')
src_cache = {}
def src(loc):
	if not loc: return ''
	f = None
	for root in ('gaussian-fluids-code_b200/csrc', '.'):
		p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), root, loc[0])
		if os.path.exists(p): f = p; break
	if not f: return ''
	if f not in src_cache: src_cache[f] = open(f).read().splitlines()
	return src_cache[f][loc[1] - 1].strip()[:110] if loc[1] <= len(src_cache[f]) else ''
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
	print(f'{100 * a[1] / tot_i:5.1f}% inst {100 * a[0] / max(tot_s, 1):5.1f}% stall  thr/inst {a[2] / max(a[1], 1):4.1f}  sass {a[3]:4d}  {loc[0] if loc else "?"}:{loc[1] if loc else 0}  {src(loc)}')
