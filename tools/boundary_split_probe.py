"""development probe: boundary forward / adjoint + gather at 8192 and 4096 samples (S1 field), each class 20 calls per CUDA graph"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gaussian_fluids_code_b200 import gsr3d
from gaussian_fluids_code_b200.init_cond3d import sample_on_box
from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field
gsr3d.device = torch.device('cuda', 0)
P, S, R, V, mgs, _ = synthetic_field(10)
o = make_fast3d(P, S, R, V, 5e-3, mgs)
e = o._engine
e.ensure_packed(o._params())
for Qb in (8192, 4096, 2048):
	bdata, bnormal = sample_on_box(Qb, 0., 1., 0., 1., 0., 1.)
	bins = e.bin_samples(bdata, True, tag='b%d' % Qb)
	valb = torch.empty((Qb, 3), device='cuda')
	f = bench.timeit_graph(lambda: e.forward(bdata, valb, None, accumulate=False, perm=bins))
	g = bench.timeit_graph(lambda: e.backward_gather(bdata, bins.perm, bins.scs, valb, None, (0., 10., 0., 0., 0., 0.), {'normals': bnormal}, None, tag='acc_b%d' % Qb, want_losses=True))
	b = bench.timeit_graph(lambda: e.bin_samples(bdata, True, tag='b%d' % Qb))
	print(f'Qb={Qb}: forward {f * 1e3:.2f} us, adjoint + gather {g * 1e3:.2f} us, bin {b * 1e3:.2f} us', flush=True)
