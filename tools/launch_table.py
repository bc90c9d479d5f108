"""aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_table.py launches.csv [out.csv]"""
import csv, re, sys
rows = []
with open(sys.argv[1]) as f:
	lines = [l for l in f if not l.startswith('==')]
for row in csv.DictReader(lines):
	try:
		ns = float(row['Metric Value'].replace(',', ''))
	except ValueError:
		ns = float('nan')
	if ns != ns:
		continue	# a launch ncu could not time
	rows.append((row['Kernel Name'], ns, row['Grid Size'], row['Block Size']))
agg = {}
for nm, ns, grid, block in rows:
	key = re.sub(r'\(.*', '', nm).replace('void ', '')
	key = key if len(key) < 110 else key[:50] + '...' + key[-55:]
	a = agg.setdefault(key, [0, 0.])
	a[0] += 1
	a[1] += ns
tot = sum(v[1] for v in agg.values())
out = [('kernel', 'launches', 'total_us', 'mean_us', 'share')]
for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
	out.append((k, c, f'{ns / 1e3:.1f}', f'{ns / c / 1e3:.2f}', f'{ns / tot:.4f}'))
out.append(('TOTAL', len(rows), f'{tot / 1e3:.1f}', '', '1'))
w = csv.writer(open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout)
w.writerows(out)
