import ctypes as C, os, sys, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from gaussian_fluids_code_b200 import gsr3d, _lib
from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field
gsr3d.device = torch.device('cuda', 0)
def timeit(fn, reps=10, warm=3):
	for _ in range(warm): fn()
	torch.cuda.synchronize()
	evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
	for a, b in evs:
		a.record(); fn(); b.record()
	torch.cuda.synchronize()
	return float(np.median([a.elapsed_time(b) for a, b in evs]))
for n in (10, 20, 40):
	P, S, R, V, mgs, gen = synthetic_field(n)
	o = make_fast3d(P, S, R, V, 5e-3, mgs)
	e = o._engine
	x = gsr3d.get_grid_points(0., 1., 0., 1., 0., 1., 128, 128, 128).contiguous()
	e.ensure_packed(o._params())
	bins = e.bin_samples(x, True)
	rv = torch.zeros((x.shape[0], 3), device='cuda'); rh = torch.zeros((x.shape[0],), device='cuda')
	out = {}
	for spc, cap in ((1 << 30, 128), (1, 128), (1, 256), (1, 64)):
		_lib.lib().gsr_set_tuning(C.c_int(12), C.c_int(spc)); _lib.lib().gsr_set_tuning(C.c_int(11), C.c_int(cap))
		out[f'spc>={spc if spc < 1e9 else "inf"},cap{cap}'] = round(timeit(lambda: e.advected_vorticity(x, -.02, rv, rh, perm=bins)), 3)
	print('n', n, 'samples/cell', round(x.shape[0] / e.ncell, 1), out, flush=True)
