"""kernel timeline of a few captured project iterations (development tool): python tools/timeline_probe.py [n] [out.csv]
CUPTI activity records through torch.profiler: start, duration and stream of every kernel of one frame; prints the iterations'
period, the busy time of the critical chain and the gaps between its kernels for a window in the middle of the frame."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from gaussian_fluids_code_b200 import timestep3d, gsr3d

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
out = sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/timeline.csv'
gsr3d.device = torch.device('cuda', 0)
ts = timestep3d.LeapfrogTimestep(n=n, iters=200, check_iter=100)
for _ in range(3):
	ts.reset()
	ts.step()
torch.cuda.synchronize()
ts.reset()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
	ts.step()
	torch.cuda.synchronize()
ev = []
for e in prof.events():
	if e.device_type == torch.autograd.DeviceType.CUDA and e.name and 'Memcpy' not in e.name and 'Memset' not in e.name:
		ev.append((e.time_range.start, e.time_range.end - e.time_range.start, e.name))
ev.sort()
t0 = ev[0][0]
os.makedirs(os.path.dirname(out) or '.', exist_ok=True)
with open(out, 'w') as f:
	f.write('start_us,dur_us,kernel\n')
	for s, d, nm in ev:
		f.write(f'{s - t0:.2f},{d:.2f},"{nm[:90]}"\n')
steps = [(s - t0, d) for s, d, nm in ev if 'step_cluster4' in nm]
print('kernels', len(ev), 'step launches', len(steps))
if len(steps) > 60:
	mid = steps[40:60]
	per = [(b[0] - a[0]) for a, b in zip(mid[:-1], mid[1:])]
	print('iteration period us: mean %.2f min %.2f max %.2f' % (sum(per) / len(per), min(per), max(per)))
	a, b = mid[0][0], mid[3][0]
	print('--- three iterations ---')
	for s, d, nm in ev:
		if a <= s - t0 < b:
			print('%9.2f %7.2f  %s' % (s - t0 - a, d, nm[:100]))
