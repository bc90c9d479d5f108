"""development probe: per-step times of the timed and the end-to-end loops of bench.py.  python tools/e2e_probe.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussian_fluids_code_b200 import gsr3d, timestep3d
gsr3d.device = dev = torch.device('cuda', 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ts = timestep3d.LeapfrogTimestep(n=n, iters=600, test_res=128)
for _ in range(3):
	ts.reset(); ts.step()
torch.cuda.synchronize()
host_params = [torch.as_tensor(a).clone().pin_memory() for a in ts.params0]
host_fields = [torch.empty(ts.lattice.shape[0], dtype=torch.float32).pin_memory() for _ in range(2)]
for mode in ('async', 'sync-after', 'host-params', 'host-params+d2h'):
	times = []
	for k in range(4):
		torch.cuda.synchronize(); t0 = time.perf_counter()
		if mode.startswith('host-params'):
			ts.reset([p.to(dev, non_blocking=True) for p in host_params])
		else:
			ts.reset()
		t1 = time.perf_counter()
		vor, div = ts.step()
		t2 = time.perf_counter()
		if mode.endswith('d2h'):
			host_fields[0].copy_(vor, non_blocking=True); host_fields[1].copy_(div, non_blocking=True)
		if mode != 'async':
			torch.cuda.synchronize()
		t3 = time.perf_counter()
		times.append((round((t1 - t0) * 1e3, 1), round((t2 - t1) * 1e3, 1), round((t3 - t2) * 1e3, 1)))
	torch.cuda.synchronize()
	print(mode, '(reset, step host time, tail sync) ms:', times, flush=True)
