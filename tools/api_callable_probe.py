"""development probe: one S1 frame through advance3d.advance_frame with ordinary callables as generators (the scene's own boundary sampler and a
torch.rand lambda, marked graph_safe) instead of the stock sampler objects — the path a driver that keeps its own generators takes"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussian_fluids_code_b200 import advance3d, gsr3d, init_cond3d
from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field
gsr3d.device = torch.device('cuda', 0)
P, S, R, V, mgs, _ = synthetic_field(10)
box = (0., 1.) * 3
test_gen = advance3d.LatticeGenerator(*box, 128, 128, 128)
bgen = init_cond3d.make_boundary_sampler('leapfrog')
for mode in ('eager', 'graph_safe', 'stock'):
	a, b = make_fast3d(P, S, R, V, 5e-3, mgs), make_fast3d(P, S, R, V, 5e-3, mgs)
	dgen = lambda n, gv: torch.rand_like(gv.positions.detach())
	bg = bgen
	if mode == 'graph_safe':
		dgen.graph_safe = True
	if mode == 'eager':
		bg = lambda n: bgen(n)
	if mode == 'stock':
		dgen, bg = advance3d.BoxSampler(*box), advance3d.BoxSurfaceSampler(*box)
	ts = []
	for k in range(5):
		torch.cuda.synchronize(); t0 = time.perf_counter()
		a, b, ep, _ = advance3d.advance_frame(a, b, *box, .02, dgen, test_gen, boundary_generator=bg, boundary_lambda=10., max_epoch=600, patience=10 ** 9, verbose=0, batch_size=8192)
		torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
	print(mode, [round(t, 1) for t in ts], flush=True)
