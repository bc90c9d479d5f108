"""development probe (needs a build with GSR_NVCC_EXTRA=-DGSR_STEP_TIMING): phase timestamps inside step_cluster4_kernel.
python tools/step_timing_probe.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussian_fluids_code_b200 import gsr3d, timestep3d, _lib
gsr3d.device = torch.device('cuda', 0)
ts = timestep3d.LeapfrogTimestep(n=10, iters=40, test_res=32, check_iter=20, use_graph=False)
ts.step()
torch.cuda.synchronize()
lib = _lib.lib()
acc = []
for _ in range(5):
	ts.reset(); ts.step(); torch.cuda.synchronize()
	out = (C.c_ulonglong * 16)()
	assert lib.gsr_debug_step_stamps(out) == 0
	t = np.array(list(out)[:12], dtype=np.int64)
	acc.append(t - t[0])
names = ['start', 'chain rule done', 'before sync1', 'after sync1', 'tail done', 'adam done', 'after sync2', 'before sync3 (keys)', 'after sync3', 'scan done', 'pack done', 'after final sync']
med = np.median(np.array(acc), axis=0)
for n, v, d in zip(names, med, np.diff(np.concatenate([[0], med]))):
	print(f'{n:24s} {v / 1e3:7.2f} us  (+{d / 1e3:.2f})')
