"""development probe: device time of one S1 time step for a given number of iterations per captured graph (GSR_GRAPH_UNIT):
600 replays of 1 iteration against 60 replays of 10 shows what a replay costs (~5 us).  (step() ends with a host read of
grid_scale, so the host time printed beside it always equals the device time and says nothing about launch-boundness.)"""
import sys, time
sys.path.insert(0, '.')
import torch
from gaussian_fluids_code_b200 import timestep3d, gsr3d
gsr3d.device = torch.device('cuda', 0)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 600
ts = timestep3d.LeapfrogTimestep(n=10, iters=iters, test_res=int(sys.argv[2]) if len(sys.argv) > 2 else 128, check_iter=100)
for _ in range(3):
	ts.reset(); ts.step()
torch.cuda.synchronize()
res = []
for _ in range(4):
	ts.reset()
	torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	t0 = time.perf_counter(); e0.record()
	ts.step()
	e1.record(); t1 = time.perf_counter()
	torch.cuda.synchronize(); t2 = time.perf_counter()
	res.append((e0.elapsed_time(e1), (t1 - t0) * 1e3, (t2 - t0) * 1e3))
import os
print({'unit': os.environ.get('GSR_GRAPH_UNIT', '1'), 'iters': iters, 'device_ms': [round(r[0], 2) for r in res], 'host_enqueue_ms': [round(r[1], 2) for r in res], 'wall_ms': [round(r[2], 2) for r in res]})
