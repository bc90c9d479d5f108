"""full-size density advection: 512^3 lattice, two fields, synthetic field of n^3 Gaussians. python tools/density_probe.py [n] [res]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from gaussian_fluids_code_b200 import advance_density, gsr3d, init_cond3d, _lib
from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
res = int(sys.argv[2]) if len(sys.argv) > 2 else 512
gsr3d.device = torch.device('cuda', 0)
P, S, R, V, mgs, gen = synthetic_field(n)
gv = make_fast3d(P, S, R, V, 5e-3, mgs)
dom = (0., 1., 0., 1., 0., 1.)
adv = advance_density.DensityAdvector(*dom, res=(res,) * 3)
info = init_cond3d.other_info['ring_collide']
d1, d2 = adv.set_ring(info['ring1']), adv.set_ring(info['ring2'])
# work census on a random 1/64 subsample of the lattice
sub = torch.rand((res ** 3 // 64, 3), device='cuda')
cnt = torch.zeros(2, dtype=torch.int64, device='cuda')
gv._engine.ensure_packed(gv._params())
gv._engine.count_pairs(sub, cnt, 1, True)
Cc, Pp = [int(v) * 64 for v in cnt.tolist()]
ms = []
for _ in range(4):
	a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	a.record()
	d1, d2 = adv.advect(gv, .02, d1, d2)
	b.record()
	torch.cuda.synchronize()
	ms.append(a.elapsed_time(b))
t = float(np.median(ms[1:]))
fma = C.c_double(0.)
_lib.lib().gsr_peak_fma(C.c_int(20000), C.byref(fma), _lib.stream())
flop = 4 * (24 * Cc + 7 * Pp)	# 4 evaluations, u only
print(json.dumps({'n': n, 'N': n ** 3, 'lattice': res, 'voxels': res ** 3, 'ms_per_frame_two_fields': t, 'all_ms': ms, 'candidate_visits_per_frame': 4 * Cc, 'pair_evals_per_s': 4 * Cc / t * 1e3,
				  'TFLOPs': flop / t / 1e9, 'frac_of_ffma_peak': flop / t / 1e9 / fma.value, 'density_mass': [float(d1.sum()), float(d2.sum())]}))
