"""
GPU tests of the per-iteration optimisation step: the fused device-resident iteration against
the reference-structured one (torch autograd regularisers, host-side PCGrad, torch.optim.Adam, ReduceLROnPlateau) fed
with the same injected samples.  Tolerance (north_star): 1e-4 relative on the per-step trajectories.
"""
import numpy as np
import pytest
import torch

from helpers import NAMES, rel_err
from test_gpu_3d_kernels import make_fast3d, synthetic

pytestmark = pytest.mark.gpu


def fields(n=8):
	P, S, R, V, mgs, gen = synthetic(n)
	return make_fast3d(P, S, R, V, 5e-3, mgs), make_fast3d(P, S, R, V, 5e-3, mgs), gen


def test_advected_vorticity_matches_rk4_composition():
	"""fused a5 kernel == RK4 (a4) followed by the reference's torch glue (3D/advance.py:35-47)"""
	from gaussian_fluids_code_b200.advance3d import curl
	cur, _, gen = fields(10)
	x = torch.rand((4000, 3), generator=gen).cuda()
	vor, hel = cur.advected_vorticity(x, .02, need_hel=True)
	psi, dpsi, pb_v, pb_dv = cur.advection_rk4(x, -.02, pos_only=False)
	pb_vor = curl(pb_dv)
	ref_hel = (pb_v * pb_vor).sum(dim=-1)
	ref_vor = (dpsi.inverse() @ pb_vor.unsqueeze(-1)).squeeze(-1)
	assert rel_err(vor.cpu().numpy(), ref_vor.cpu().numpy()) < 1e-5
	assert rel_err(hel.cpu().numpy(), ref_hel.cpu().numpy()) < 1e-5


@pytest.mark.parametrize('boundary_lambda', [0., 10.])
def test_fused_iteration_matches_unfused(boundary_lambda):
	from gaussian_fluids_code_b200 import advance3d
	from gaussian_fluids_code_b200.init_cond3d import sample_on_box
	iters = 3
	results = {}
	for fused in (False, True):
		cur, new, gen = fields(8)
		N = new.N
		datas = [torch.rand((N, 3), generator=gen).cuda() for _ in range(iters)]
		torch.manual_seed(7)
		bnds = [sample_on_box(2048, 0., 1., 0., 1., 0., 1.) for _ in range(iters)]
		before = [getattr(new, nm).detach().clone() for nm in NAMES]
		ref = advance3d.AdvectedCovectorField(cur, cur, .02, 0., 1., 0., 1., 0., 1.)
		it_d, it_b = iter(datas), iter(bnds)
		advance3d.project(new, ref, 0., 1., 0., 1., 0., 1., lambda n, gv: next(it_d), lambda gv: datas[0],
						  boundary_generator=(lambda n: next(it_b)) if boundary_lambda else None, boundary_lambda=boundary_lambda,
						  max_epoch=iters, verbose=0, fused=fused, check_iter=1000)
		results[fused] = ([getattr(new, nm).detach().cpu().numpy() for nm in NAMES], [b.cpu().numpy() for b in before], new.grid_scale)
	(pa, b0, gs_a), (pb, _, gs_b) = results[False], results[True]
	assert np.float32(gs_a) == np.float32(gs_b)
	for nm, a, b, b_ in zip(NAMES, pa, pb, b0):
		assert rel_err(b, a) < 1e-4, nm			# trajectory tolerance of the north star
		da, db = a - b_, b - b_				# and, much stricter, the parameter UPDATES themselves
		assert np.abs(da).max() > 0
		assert rel_err(db, da) < 2e-2, (nm, rel_err(db, da))


def test_fused_project_decreases_losses_and_stops():
	"""a short real run: losses go down, the scheduler state lives on device, early-stop bookkeeping runs"""
	from gaussian_fluids_code_b200 import advance3d
	from gaussian_fluids_code_b200.init_cond3d import sample_on_box
	cur, new, gen = fields(8)
	torch.manual_seed(3)
	advance3d.advect_covector_field(new, cur, .02, new.x_min, new.x_max, new.y_min, new.y_max, new.z_min, new.z_max)
	ref = advance3d.AdvectedCovectorField(cur, cur, .02, 0., 1., 0., 1., 0., 1.)
	test_pts = torch.rand((20000, 3), generator=gen).cuda()
	hist = {}
	ep = advance3d.project(new, ref, 0., 1., 0., 1., 0., 1., lambda n, gv: torch.rand_like(gv.positions), lambda gv: test_pts,
						   boundary_generator=lambda n: sample_on_box(n, 0., 1., 0., 1., 0., 1.), boundary_lambda=10., batch_size=2048,
						   max_epoch=300, patience=200, verbose=0, check_iter=50, history=hist)
	assert ep <= 300 and len(hist['test']) >= 1
	first, last = hist['test'][0], hist['test'][-1]
	assert last['loss_div'] <= first['loss_div'] * 1.5
	assert all(np.isfinite(list(t.values())).all() for t in hist['test'])
	# the generic API works again after the fused phase
	val, grad = new.get_losses(test_pts[:100].contiguous())
	assert torch.isfinite(val).all() and torch.isfinite(grad).all()


def test_fused_fit_matches_unfused_3d():
	"""3D/initialize.py:9-46: value + gradient L1 fit, fused iteration vs the reference-structured one on injected samples"""
	from gaussian_fluids_code_b200 import init_cond3d, initialize3d
	field = init_cond3d.make_field('leapfrog')
	res = {}
	for fused in (False, True):
		_, gv, gen = fields(8)
		datas = iter([torch.rand((gv.N, 3), generator=gen).cuda() for _ in range(3)])
		before = [getattr(gv, nm).detach().cpu().numpy().copy() for nm in NAMES]
		initialize3d.fit_velocity_with_gradient(gv, field, field.gradient, lambda n: next(datas), max_epoch=3, verbose=0, fused=fused)
		res[fused] = ([getattr(gv, nm).detach().cpu().numpy() for nm in NAMES], before)
	for nm, a, b, b_ in zip(NAMES, res[False][0], res[True][0], res[False][1]):
		assert rel_err(b, a) < 1e-4, nm
		assert rel_err(b - b_, a - b_) < 2e-2, (nm, rel_err(b - b_, a - b_))


def test_simulation_initialize_fits_the_leapfrog_rings():
	"""a short fit of the repo's own 3D leapfrog scene (10^3 Gaussians): the value loss must go down substantially"""
	from gaussian_fluids_code_b200 import gsr3d, init_cond3d, initialize3d
	gsr3d.device = torch.device('cuda', 0)
	torch.manual_seed(0)
	field = init_cond3d.make_field('leapfrog')
	x = torch.rand((20000, 3), device='cuda')
	ref = field(x)
	gv = initialize3d.simulation_initialize('leapfrog', max_epoch=150, verbose=0)
	err = float((gv(x) - ref).abs().mean() / ref.abs().mean())
	assert np.isfinite(err) and err < .8, err	# from 1.0 (zero field) after 150 of the reference's 500 epochs


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('epochs', [1, 3])
def test_project_matches_reference_golden(fused, epochs):
	"""the per-timestep optimisation against the reference's OWN project() (3D/advance.py:183-334: PCGrad, autograd regularisers,
	4 x Adam, 4 x ReduceLROnPlateau, grid rebuild) run on its own GaussianSplatting3DFast through the Taichi shim with recorded sample
	batches (tests/golden/make_golden_project3d.py): the parameter updates after 1 and 3 iterations, the next grid_scale, the lrs"""
	from helpers import load_golden
	from gaussian_fluids_code_b200 import advance3d, gsr3d
	gsr3d.device = torch.device('cuda', 0)
	g = load_golden('ref3d_project.npz')

	def field(P):
		gv = gsr3d.GaussianSplatting3DFast(0., 1., 0., 1., 0., 1., P, dim=3)	# the reference's defaults: min_grid_scale and clamp_threshold
		assert gv.min_grid_scale == pytest.approx(float(g['min_grid_scale']), rel=1e-12) and gv.clamp_threshold == float(g['tau'])
		with torch.no_grad():
			gv.scalings.copy_(torch.tensor(g['scalings'])); gv.rotations.copy_(torch.tensor(g['rotations'])); gv.values.copy_(torch.tensor(g['values']))
		gv.reinitialize_grid()
		gv.zero_grad()
		return gv
	cur, new = field(g['cur_positions']), field(g['new_positions'])
	ref = advance3d.AdvectedCovectorField(cur, cur, float(g['dt']), 0., 1., 0., 1., 0., 1.)
	datas = iter([torch.tensor(x, device='cuda') for x in g['samples']])
	bnds = iter([(torch.tensor(d, device='cuda'), torch.tensor(n, device='cuda')) for d, n in zip(g['boundary_data'], g['boundary_normal'])])
	advance3d.project(new, ref, 0., 1., 0., 1., 0., 1., lambda n, gv: next(datas), lambda gv: None, boundary_generator=lambda n: next(bnds),
					  boundary_lambda=float(g['boundary_lambda']), batch_size=g['boundary_data'].shape[1], max_epoch=epochs, patience=500, verbose=0,
					  fused=fused, check_iter=1000)
	before = dict(positions=g['new_positions'], scalings=g['scalings'], rotations=g['rotations'], values=g['values'])
	for nm in NAMES:
		got = getattr(new, nm).detach().cpu().numpy()
		want = g[f'after{epochs}_{nm}']
		d_ref, d_got = want - before[nm], got - before[nm]
		assert np.abs(d_ref).max() > 0
		assert rel_err(got, want) < 1e-5, nm	# trajectory
		assert rel_err(d_got, d_ref) < 2e-2, (nm, rel_err(d_got, d_ref))	# the updates themselves (same bound as fused vs unfused)
	assert new.grid_scale == pytest.approx(float(g[f'after{epochs}_grid_scale']), rel=2e-6)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('epochs', [1, 3])
def test_fit_matches_reference_golden(fused, epochs):
	"""the initial fit against the reference's OWN fit_velocity_with_gradient (3D/initialize.py:9-46) run on its own
	GaussianSplatting3DFast through the Taichi shim with recorded batches and targets (tests/golden/make_golden_fit3d.py)"""
	from helpers import load_golden
	from gaussian_fluids_code_b200 import gsr3d, initialize3d
	gsr3d.device = torch.device('cuda', 0)
	g = load_golden('ref3d_fit.npz')
	gv = gsr3d.GaussianSplatting3DFast(0., 1., 0., 1., 0., 1., g['positions'], dim=3)
	np.testing.assert_allclose([gv.positions_lr, gv.scalings_lr, gv.rotations_lr, gv.values_lr], g['lrs'])	# the class's default learning rates
	with torch.no_grad():
		gv.scalings.copy_(torch.tensor(g['scalings'])); gv.rotations.copy_(torch.tensor(g['rotations'])); gv.values.copy_(torch.tensor(g['values']))
	gv.reinitialize_grid()
	gv.zero_grad()
	T = lambda a: torch.tensor(a, device='cuda')
	datas, vals, grads = iter([T(x) for x in g['samples']]), iter([T(x) for x in g['ref_val']]), iter([T(x) for x in g['ref_grad']])
	initialize3d.fit_velocity_with_gradient(gv, lambda x: next(vals), lambda x: next(grads), lambda n: next(datas), batch_size=g['samples'].shape[1],
											max_epoch=epochs, verbose=0, fused=fused)
	for nm in NAMES:
		got, want = getattr(gv, nm).detach().cpu().numpy(), g[f'after{epochs}_{nm}']
		d_ref, d_got = want - g[nm], got - g[nm]
		assert np.abs(d_ref).max() > 0
		assert rel_err(got, want) < 1e-5, nm
		assert rel_err(d_got, d_ref) < 2e-2, (nm, rel_err(d_got, d_ref))
	assert gv.grid_scale == pytest.approx(float(g[f'after{epochs}_grid_scale']), rel=2e-6)
