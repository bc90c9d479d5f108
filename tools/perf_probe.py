"""quick kernel timing probe (development tool): python tools/perf_probe.py n [Q] [tiled_min_q] [cap]"""
import sys, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np, torch
from gaussian_fluids_code_b200 import gsr3d, _lib
from gaussian_fluids_code_b200.engine import HashEngine
from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field

def timeit(fn, reps=10, warm=3):
	for _ in range(warm): fn()
	torch.cuda.synchronize()
	evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
	for a, b in evs:
		a.record(); fn(); b.record()
	torch.cuda.synchronize()
	return float(np.median([a.elapsed_time(b) for a, b in evs]))

n = int(sys.argv[1]); N = n ** 3
Q = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else N
if len(sys.argv) > 3: HashEngine.set_tiled_min_q(int(sys.argv[3]))
if len(sys.argv) > 4 and int(sys.argv[4]) >= 0: _lib.check(_lib.lib().gsr_set_tuning(C.c_int(2), C.c_int(int(sys.argv[4]))), 'cap')
if len(sys.argv) > 5: _lib.check(_lib.lib().gsr_set_tuning(C.c_int(4), C.c_int(int(sys.argv[5]))), 'fwp4')
gsr3d.device = torch.device('cuda', 0)
P, S, R, V, mgs, gen = synthetic_field(n)
o = make_fast3d(P, S, R, V, 5e-3, mgs)
e = o._engine
lattice = os.environ.get('LATTICE')
if lattice:
	r = int(lattice); x = gsr3d.get_grid_points(0., 1., 0., 1., 0., 1., r, r, r).contiguous(); Q = x.shape[0]
else:
	x = torch.rand((Q, 3), generator=gen).cuda()
e.ensure_packed(o._params())
bins = e.bin_samples(x, True)
cnt = torch.zeros(2, dtype=torch.int64, device='cuda')
e.count_pairs(x, cnt, 1, True)
Cc, Pp = [int(v) for v in cnt.tolist()]
res = {'n': n, 'N': N, 'Q': Q, 'tiled': bins.tiles is not None, 'C_per_Q': Cc / Q, 'P_per_Q': Pp / Q}
flop = 24 * Cc + 28 * Pp
val = torch.zeros((Q, 3), device='cuda'); grad = torch.zeros((Q, 3, 3), device='cuda')
t = timeit(lambda: e.forward(x, val, grad, False, bins)); res['fwd_ms'] = t; res['fwd_Gcand_s'] = Cc / t / 1e6; res['fwd_TFLOPs'] = flop / t / 1e9
t = timeit(lambda: e.forward(x, val, None, False, bins)); res['fwd_valonly_ms'] = t
rv = torch.zeros((Q, 3), device='cuda'); rh = torch.zeros((Q,), device='cuda')
t = timeit(lambda: e.advected_vorticity(x, -.02, rv, rh, perm=bins)); res['rk4_pullback_ms'] = t; res['rk4_Gcand_s'] = 5 * Cc / t / 1e6; res['rk4_TFLOPs'] = 5 * flop / t / 1e9
gp = torch.zeros((Q, 3), device='cuda')
t = timeit(lambda: e.rk4(x, -.02, gp)); res['rk4_posonly_incl_binning_ms'] = t
ref_vor = torch.randn((Q, 3), device='cuda') * .1; ref_hel = torch.randn((Q,), device='cuda') * .1
e.forward(x, val, grad, False, bins)
if Q <= 4 * N or os.environ.get('BWD'):
	t = timeit(lambda: e.backward_gather(x, bins.perm, bins.scs, val, grad, (0, 0, 0, 1., 1., 1.), {'ref_vor': ref_vor, 'ref_hel': ref_hel}, None)); res['bwd_gather_ms'] = t; res['bwd_Gcand_s'] = Cc / t / 1e6
	res['bwd_TFLOPs'] = (24 * Cc + 230 * Pp) / t / 1e9
t = timeit(lambda: e.bin_samples(x, True)); res['bin_samples_ms'] = t
t = timeit(lambda: e.build(o.positions.detach())); res['build_grid_ms'] = t
t = timeit(lambda: (setattr(e, '_packed_key', None), e.ensure_packed(o._params()))); res['pack_ms'] = t
print(json.dumps(res))
