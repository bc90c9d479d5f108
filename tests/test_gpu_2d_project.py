"""
GPU tests of the 2D time stepper (SURVEY 8a rows a5, a7, a8 in 2D; BASELINE config 1, the Taylor-Green vortex):
the fused device-resident iteration against the reference-structured one on injected samples (1e-4 per-step trajectory
tolerance of the north star), the fused initial fit, and a physical known-answer test — Taylor-Green is a steady
solution of the Euler equations, so the fitted field must stay put when it is advanced.
"""
import numpy as np
import pytest
import torch

from helpers import NAMES, rel_err

pytestmark = pytest.mark.gpu


def scene_and_fields(n=16, seed=5):
	"""two identical random-ish 2D fields on the Taylor-Green domain (GSR space: [0,10]^2)"""
	from gaussian_fluids_code_b200 import gsr2d
	from gaussian_fluids_code_b200.init_cond2d import Scene2D
	gsr2d.device = torch.device('cuda', 0)
	sc = Scene2D('taylor_green')
	x0, x1, y0, y1 = sc.scaled(sc.initialize_domain)
	pts = gsr2d.get_grid_points(x0, x1, y0, y1, n, n).cpu().numpy()
	gen = torch.Generator().manual_seed(seed)
	out = []
	S = torch.randn((n * n, 2), generator=gen) * .08
	R = torch.rand((n * n,), generator=gen) * 6.28
	V = torch.randn((n * n, 2), generator=gen) * .3
	for _ in range(2):
		o = gsr2d.GaussianSplattingFast(x0, x1, y0, y1, pts, dim=2)
		with torch.no_grad():
			o.scalings += S.cuda()
			o.rotations.copy_(R.cuda())
			o.values.copy_(V.cuda())
		o.zero_grad()
		out.append(o)
	return sc, out[0], out[1], gen


@pytest.mark.parametrize('boundary_lambda', [0., 1.])
def test_fused_iteration_matches_unfused_2d(boundary_lambda):
	from gaussian_fluids_code_b200 import advance2d
	iters = 3
	results = {}
	for fused in (False, True):
		sc, cur, new, gen = scene_and_fields()
		N = new.N
		datas = [(torch.rand((N, 2), generator=gen) * 10.).cuda() for _ in range(iters)]
		torch.manual_seed(11)
		bnds = [sc.boundary_sampler_2(512) for _ in range(iters)]
		before = [getattr(new, nm).detach().clone() for nm in NAMES]
		ref = advance2d.AdvectedCovectorField(cur, cur, .01, domain=sc.scaled(sc.advance_domain))
		it_d, it_b = iter(datas), iter(bnds)
		advance2d.project(new, ref, lambda n, gv: next(it_d), lambda gv: datas[0], boundary_generator_2=(lambda n: next(it_b)) if boundary_lambda else None,
						  boundary_lambda=boundary_lambda, max_epoch=iters, verbose=0, fused=fused, check_iter=1000)
		results[fused] = ([getattr(new, nm).detach().cpu().numpy() for nm in NAMES], [b.cpu().numpy() for b in before], new.grid_scale)
	(pa, b0, gs_a), (pb, _, gs_b) = results[False], results[True]
	assert np.float32(gs_a) == np.float32(gs_b)
	for nm, a, b, b_ in zip(NAMES, pa, pb, b0):
		assert rel_err(b, a) < 1e-4, nm
		da, db = a - b_, b - b_
		assert np.abs(da).max() > 0
		assert rel_err(db, da) < 2e-2, (nm, rel_err(db, da))


def test_fused_fit_matches_unfused_2d():
	from gaussian_fluids_code_b200 import advance2d
	res = {}
	for fused in (False, True):
		sc, gv, _, gen = scene_and_fields(seed=9)
		datas = iter([(torch.rand((gv.N, 2), generator=gen) * 10.).cuda() for _ in range(3)])
		gv.set_lr(positions_lr=1.6e-3, scalings_lr=5e-2, rotations_lr=5e-2, values_lr=5e-3)
		before = [getattr(gv, nm).detach().cpu().numpy().copy() for nm in NAMES]
		advance2d.fit_velocity_with_gradient(gv, sc.target_velocity, sc.target_gradient, lambda n: next(datas), max_epoch=3, verbose=0, fused=fused)
		res[fused] = ([getattr(gv, nm).detach().cpu().numpy() for nm in NAMES], before)
	for nm, a, b, b_ in zip(NAMES, res[False][0], res[True][0], res[False][1]):
		assert rel_err(b, a) < 1e-4, nm
		assert rel_err(b - b_, a - b_) < 2e-2, (nm, rel_err(b - b_, a - b_))


def test_taylor_green_is_steady():
	"""BASELINE config 1 in miniature: fit the 24 x 24 Taylor-Green field, advance it a few steps of dt = .001 (advect +
	project), and compare with the analytic steady solution (2D/init_cond.py:158-167)."""
	from gaussian_fluids_code_b200 import advance2d, gsr2d
	from gaussian_fluids_code_b200.init_cond2d import Scene2D
	gsr2d.device = torch.device('cuda', 0)
	torch.manual_seed(42)
	sc = Scene2D('taylor_green')
	gv = advance2d.simulation_initialize(sc, max_epoch=1500, verbose=0)
	x = sc.test_generator()
	ref = sc.target_velocity(x)
	err0 = float((gv(x) - ref).abs().mean() / ref.abs().mean())
	assert err0 < .05, err0		# the fit itself (10 000 epochs in the reference; 1 500 here)
	x0, x1, y0, y1 = sc.scaled(sc.initialize_domain)
	spare = gsr2d.GaussianSplattingFast(x0, x1, y0, y1, gv.positions.detach().cpu().numpy(), dim=2)
	cur = gv
	for _ in range(3):
		cur, spare = advance2d.advance(sc, cur, spare, .001 * sc.scaling_factor ** 0, max_epoch=200, verbose=0)
	err1 = float((cur(x) - ref).abs().mean() / ref.abs().mean())
	assert np.isfinite(err1) and err1 < err0 + .02, (err0, err1)
	g = cur.gradient(x)
	div = (g[:, 0, 0] + g[:, 1, 1]).abs().mean()
	vor = (g[:, 1, 0] - g[:, 0, 1]).abs().mean()
	assert float(div / vor) < .1


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('epochs', [1, 3])
def test_project_matches_reference_golden(fused, epochs):
	"""the 2D per-timestep optimisation against the reference's OWN project() (2D/advance.py:186-291: value samples on obstacles,
	normal samples, PCGrad, autograd regularisers with the position-drift term, 4 x Adam, 4 x ReduceLROnPlateau, grid rebuild) run on
	its own GaussianSplattingFast through the Taichi shim with recorded batches (tests/golden/make_golden_project2d.py)"""
	from helpers import load_golden, rel_err
	from gaussian_fluids_code_b200 import advance2d, gsr2d
	gsr2d.device = torch.device('cuda', 0)
	g = load_golden('ref2d_project.npz')
	dom = tuple(float(v) for v in g['domain'])

	def field(P):
		gv = gsr2d.GaussianSplattingFast(*dom, P, dim=2)
		assert gv.min_grid_scale == pytest.approx(float(g['min_grid_scale']), rel=1e-12) and gv.clamp_threshold == float(g['tau'])
		with torch.no_grad():
			gv.scalings.copy_(torch.tensor(g['scalings'])); gv.rotations.copy_(torch.tensor(g['rotations'])); gv.values.copy_(torch.tensor(g['values']))
		gv.reinitialize_grid()
		gv.zero_grad()
		return gv
	cur, new = field(g['cur_positions']), field(g['new_positions'])
	ref = advance2d.AdvectedCovectorField(cur, cur, float(g['dt']), domain=dom)	# taylor_vortex: scale factor 1
	T = lambda a: torch.tensor(a, device='cuda')
	datas = iter([T(x) for x in g['samples']])
	g1 = iter([(T(d), T(v)) for d, v in zip(g['b1_data'], g['b1_val'])])
	g2 = iter([(T(d), T(n), T(r)) for d, n, r in zip(g['b2_data'], g['b2_normal'], g['b2_ref'])])
	advance2d.project(new, ref, lambda n, gv: next(datas), lambda gv: None, boundary_generator_1=lambda n: next(g1), boundary_generator_2=lambda n: next(g2),
					  boundary_lambda=float(g['boundary_lambda']), batch_size=g['b1_data'].shape[1], max_epoch=epochs, patience=500, verbose=0, fused=fused)
	before = dict(positions=g['new_positions'], scalings=g['scalings'], rotations=g['rotations'], values=g['values'])
	for nm in ('positions', 'scalings', 'rotations', 'values'):
		got = getattr(new, nm).detach().cpu().numpy().reshape(before[nm].shape)
		want = g[f'after{epochs}_{nm}']
		d_ref, d_got = want - before[nm], got - before[nm]
		assert np.abs(d_ref).max() > 0
		assert rel_err(got, want) < 1e-5, nm
		assert rel_err(d_got, d_ref) < 2e-2, (nm, rel_err(d_got, d_ref))
	assert new.grid_scale == pytest.approx(float(g[f'after{epochs}_grid_scale']), rel=2e-6)
