"""
N2: the analytic initial fields of the 3D scenes — regularised Biot-Savart sums over ring particles
(3D/init_cond.py:122-216) — on the CUDA kernel gsr_vortex_particles, against (i) golden vectors produced by the reference's
own Taichi kernel bodies (tests/golden/make_golden_init3d.py) and (ii) the float64 oracle on seeded points.
Tolerance: 1e-5 of the largest entry per tensor (float32 arithmetic as in the reference, float64 truth).
"""
import numpy as np
import pytest
import torch

from helpers import load_golden, ring_particles_np

pytestmark = pytest.mark.gpu
SCENES = ('leapfrog', 'single_vortex_ring', 'ring_collide', 'ring_with_obstacle')
TOL = 1e-5


def maxrel(a, b):
	return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize('name', SCENES)
def test_fields_match_reference_golden(name):
	from gaussian_fluids_code_b200 import gsr3d, init_cond3d
	gsr3d.device = torch.device('cuda', 0)
	g = load_golden('ref3d_init_fields.npz')
	x = torch.tensor(g['x'], device='cuda')
	f = init_cond3d.make_field(name)
	val, grad = f(x), f.gradient(x)
	assert maxrel(val.double().cpu().numpy(), g[f'{name}_val_f64']) < TOL
	assert maxrel(grad.double().cpu().numpy(), g[f'{name}_grad_f64']) < TOL
	v2, g2 = f.both(x)	# one pass for both outputs: the same sums
	np.testing.assert_array_equal(v2.cpu().numpy(), val.cpu().numpy())
	np.testing.assert_array_equal(g2.cpu().numpy(), grad.cpu().numpy())


@pytest.mark.parametrize('name,Q', [('leapfrog', 5000), ('ring_collide', 64000), ('single_vortex_ring', 1)])
def test_fields_match_oracle_seeded(name, Q):
	import oracle.oracle as orc
	from gaussian_fluids_code_b200 import gsr3d, init_cond3d
	gsr3d.device = torch.device('cuda', 0)
	x = torch.rand((Q, 3), generator=torch.Generator().manual_seed(7))
	f = init_cond3d.make_field(name)
	val, grad = f.both(x.cuda())
	rv, rg = np.zeros((Q, 3)), np.zeros((Q, 3, 3))
	for ring in init_cond3d.rings_of(name):
		x0, w, U, a = ring_particles_np(ring, np.float32)	# the particles as the product path builds them (float32 torch ops)
		v, j = orc.vortex_particles(x.numpy(), x0.astype(np.float64), w.astype(np.float64), U, a, real=np.float64)
		rv += v
		rg += j
	assert maxrel(val.double().cpu().numpy(), rv) < TOL
	assert maxrel(grad.double().cpu().numpy(), rg) < TOL


def test_accumulates_and_rejects_cpu_tensors():
	from gaussian_fluids_code_b200 import _lib, gsr3d, init_cond3d
	gsr3d.device = torch.device('cuda', 0)
	ring = init_cond3d.rings_of('single_vortex_ring')[0]
	x = torch.rand((100, 3), generator=torch.Generator().manual_seed(1)).cuda()
	once = init_cond3d.vortex_ring(x, ring)
	twice = once.clone()
	init_cond3d._biot_savart(x, ring, twice, None)	# accumulates, like the reference kernel's `res +=`
	np.testing.assert_allclose(twice.cpu().numpy(), 2. * once.cpu().numpy(), rtol=1e-6)
	with pytest.raises(_lib.GsrError):
		init_cond3d.vortex_ring(x.cpu(), ring)


def test_initial_fit_from_cuda_graphs_equals_the_eager_fit():
	"""initialize3d.simulation_initialize (the fit of 3D/initialize.py:49-86 to the analytic ring field) replayed from CUDA graphs, ten
	iterations per graph with the next batch and its Biot-Savart targets prepared on a second stream, against the same loop run eagerly:
	same kernels, same random stream — the same bits"""
	import torch
	from gaussian_fluids_code_b200 import graphloop, gsr3d, initialize3d
	gsr3d.device = torch.device('cuda', 0)
	out = {}
	for use_graph in (False, True):
		torch.manual_seed(9)
		g0 = graphloop.GRAPH_LAUNCHES
		gv = initialize3d.simulation_initialize('leapfrog', max_epoch=45, verbose=0, use_graph=use_graph)
		assert (graphloop.GRAPH_LAUNCHES > g0) == use_graph
		out[use_graph] = [p.detach().clone() for p in gv._params()] + [torch.tensor(gv.grid_scale)]
	for a, b in zip(out[False], out[True]):
		assert torch.isfinite(a).all() and torch.equal(a.cpu(), b.cpu())
