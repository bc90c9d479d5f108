"""
GPU parity of the GRADIENTS the optimiser consumes, against goldens recorded INSIDE the reference's own loops (project() of
3D/advance.py and 2D/advance.py, fit_velocity_with_gradient of 3D/initialize.py and 2D/initialize.py, run on the reference's own
classes through the Taichi shim: tests/golden/make_golden_{project3d,project2d,fit3d,fit2d}.py).  At every step() of those loops the
generators store the raw vorticity / divergence gradient sets as the loss kernels leave them (before PCGrad), the total .grad
(projected sets + autograd regularisers + boundary passes), the scheduler metric and the learning rates in use.

The parameter-update tests (test_gpu_3d_project.py, test_gpu_2d_project.py) cannot see a wrong gradient magnitude — Adam's first
update is lr * g / (|g| + eps) ~ lr * sign(g); these tests can: 2e-5 relative on iteration 1 (f32 here against the reference's f32
run), growing with the iteration because later iterations start from parameters that already differ by rounding.
Also: the 2D initial fit against its golden (parameters after 1 and 3 iterations, fused and unfused).
"""
import numpy as np
import pytest
import torch

from helpers import NAMES, load_golden, rel_err

pytestmark = pytest.mark.gpu


def record_steps(gv, rec, set_method=None):
	"""the recorder of tests/golden/ti_shim.record_steps for the CUDA classes"""
	orig_step = gv.step

	def step(metrics):
		rec.setdefault('grads', []).append({nm: getattr(gv, nm).grad.detach().cpu().numpy().copy() for nm in NAMES})
		rec.setdefault('metric', []).append(float(metrics))
		rec.setdefault('lr', []).append([o.param_groups[0]['lr'] for o in gv.optimizers])
		return orig_step(metrics)
	gv.step = step
	if set_method:
		orig = getattr(gv, set_method)

		def losses(x, *a, **kw):
			res = orig(x, *a, **kw)
			if kw.get('vor_positions_grad') is not None:
				rec.setdefault('sets', []).append({f'{t}_{nm}': kw[f'{t}_{nm}_grad'].detach().cpu().numpy().copy() for t in ('vor', 'div') for nm in NAMES})
			return res
		setattr(gv, set_method, losses)


def check_steps(g, rec, epochs, sets=True):
	for k in range(epochs):
		tol = 2e-5 * 3 ** k
		for nm in NAMES:
			for tag in (('vor', 'div') if sets else ()):
				want = g[f'it{k + 1}_{tag}_{nm}_grad']
				got = rec['sets'][k][f'{tag}_{nm}'].reshape(want.shape)
				assert rel_err(got, want) < tol, (k, tag, nm, rel_err(got, want))
			want = g[f'it{k + 1}_total_{nm}_grad']
			got = rec['grads'][k][nm].reshape(want.shape)
			assert rel_err(got, want) < tol, (k, 'total', nm, rel_err(got, want))
		assert rec['metric'][k] == pytest.approx(float(g[f'it{k + 1}_metric']), rel=2e-5)
		np.testing.assert_allclose(rec['lr'][k], g[f'it{k + 1}_lr_used'], rtol=1e-6)


def T(a):
	return torch.tensor(a, device='cuda')


def field3(g, P):
	from gaussian_fluids_code_b200 import gsr3d
	gv = gsr3d.GaussianSplatting3DFast(0., 1., 0., 1., 0., 1., P, dim=3)
	with torch.no_grad():
		gv.scalings.copy_(T(g['scalings'])); gv.rotations.copy_(T(g['rotations'])); gv.values.copy_(T(g['values']))
	gv.reinitialize_grid()
	gv.zero_grad()
	return gv


def field2(g, P):
	from gaussian_fluids_code_b200 import gsr2d
	gv = gsr2d.GaussianSplattingFast(*[float(v) for v in g['domain']], P, dim=2)
	with torch.no_grad():
		gv.scalings.copy_(T(g['scalings'])); gv.rotations.copy_(T(g['rotations']).reshape(gv.rotations.shape)); gv.values.copy_(T(g['values']))
	gv.reinitialize_grid()
	gv.zero_grad()
	return gv


def test_project3d_gradients_match_reference():
	from gaussian_fluids_code_b200 import advance3d, gsr3d
	gsr3d.device = torch.device('cuda', 0)
	g = load_golden('ref3d_project.npz')
	cur, new = field3(g, g['cur_positions']), field3(g, g['new_positions'])
	ref = advance3d.AdvectedCovectorField(cur, cur, float(g['dt']), 0., 1., 0., 1., 0., 1.)
	datas = iter([T(x) for x in g['samples']])
	bnds = iter([(T(d), T(n)) for d, n in zip(g['boundary_data'], g['boundary_normal'])])
	rec = {}
	record_steps(new, rec, 'get_losses')
	advance3d.project(new, ref, 0., 1., 0., 1., 0., 1., lambda n, gv: next(datas), lambda gv: None, boundary_generator=lambda n: next(bnds),
					  boundary_lambda=float(g['boundary_lambda']), batch_size=g['boundary_data'].shape[1], max_epoch=3, patience=500, verbose=0, fused=False, check_iter=1000)
	check_steps(g, rec, 3)


def test_project3d_fused_metric_and_learning_rates_match_reference():
	"""the fused step forms the scheduler metric (vor + div sample losses, regulariser losses, boundary loss) and the lrs on the
	device: after every iteration they must be the reference's"""
	from gaussian_fluids_code_b200 import _lib, advance3d, gsr3d
	gsr3d.device = torch.device('cuda', 0)
	g = load_golden('ref3d_project.npz')
	cur, new = field3(g, g['cur_positions']), field3(g, g['new_positions'])
	ref = advance3d.AdvectedCovectorField(cur, cur, float(g['dt']), 0., 1., 0., 1., 0., 1.)
	new.positions_lr, new.scalings_lr, new.rotations_lr, new.values_lr = [advance3d.PROJECT_LRS[k] for k in NAMES]
	fp = advance3d.FusedProjector(new, ref, float(g['boundary_lambda']), patience=50)
	for k in range(3):
		fp.iterate(T(g['samples'][k]), (T(g['boundary_data'][k]), T(g['boundary_normal'][k])))
		sc = fp.stepper.scalars()
		assert sc[_lib.ST_LOSS_TOT] == pytest.approx(float(g[f'it{k + 1}_metric']), rel=2e-5 * 3 ** k)
		assert sc[_lib.ST_T] == k + 1
		if k + 2 <= 3:	# the lrs after this step are the ones the next step uses
			np.testing.assert_allclose(sc[_lib.ST_LR:_lib.ST_LR + 4], g[f'it{k + 2}_lr_used'], rtol=1e-6)
	fp.finish()


def test_project2d_gradients_match_reference():
	from gaussian_fluids_code_b200 import advance2d, gsr2d
	gsr2d.device = torch.device('cuda', 0)
	g = load_golden('ref2d_project.npz')
	dom = tuple(float(v) for v in g['domain'])
	cur, new = field2(g, g['cur_positions']), field2(g, g['new_positions'])
	ref = advance2d.AdvectedCovectorField(cur, cur, float(g['dt']), domain=dom)
	datas = iter([T(x) for x in g['samples']])
	g1 = iter([(T(d), T(v)) for d, v in zip(g['b1_data'], g['b1_val'])])
	g2 = iter([(T(d), T(n), T(r)) for d, n, r in zip(g['b2_data'], g['b2_normal'], g['b2_ref'])])
	rec = {}
	record_steps(new, rec, 'get_grad_losses')
	advance2d.project(new, ref, lambda n, gv: next(datas), lambda gv: None, boundary_generator_1=lambda n: next(g1), boundary_generator_2=lambda n: next(g2),
					  boundary_lambda=float(g['boundary_lambda']), batch_size=g['b1_data'].shape[1], max_epoch=3, patience=500, verbose=0, fused=False)
	check_steps(g, rec, 3)


def test_fit3d_gradients_match_reference():
	from gaussian_fluids_code_b200 import gsr3d, initialize3d
	gsr3d.device = torch.device('cuda', 0)
	g = load_golden('ref3d_fit.npz')
	gv = field3(g, g['positions'])
	datas, vals, grads = iter([T(x) for x in g['samples']]), iter([T(x) for x in g['ref_val']]), iter([T(x) for x in g['ref_grad']])
	rec = {}
	record_steps(gv, rec)
	initialize3d.fit_velocity_with_gradient(gv, lambda x: next(vals), lambda x: next(grads), lambda n: next(datas), batch_size=g['samples'].shape[1], max_epoch=3, verbose=0,
											fused=False)
	check_steps(g, rec, 3, sets=False)


def fit2d(g, epochs, fused, rec=None):
	from gaussian_fluids_code_b200 import advance2d, gsr2d
	gsr2d.device = torch.device('cuda', 0)
	gv = field2(g, g['positions'])
	lrs = [float(v) for v in g['lrs']]
	gv.set_lr(positions_lr=lrs[0], scalings_lr=lrs[1], rotations_lr=lrs[2], values_lr=lrs[3])
	datas, vals, grads = iter([T(x) for x in g['samples']]), iter([T(x) for x in g['ref_val']]), iter([T(x) for x in g['ref_grad']])
	if rec is not None:
		gv.initialize_optimizers()
		record_steps(gv, rec)
	advance2d.fit_velocity_with_gradient(gv, lambda x: next(vals), lambda x: next(grads), lambda n: next(datas), batch_size=g['samples'].shape[1], max_epoch=epochs,
										 verbose=0, fused=fused)
	return gv


def test_fit2d_gradients_match_reference():
	g = load_golden('ref2d_fit.npz')
	rec = {}
	fit2d(g, 3, False, rec)
	check_steps(g, rec, 3, sets=False)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('epochs', [1, 3])
def test_fit2d_matches_reference_golden(fused, epochs):
	"""the 2D initial fit against the reference's OWN fit_velocity_with_gradient (2D/initialize.py:10-41) run on its own
	GaussianSplattingFast through the Taichi shim with recorded batches and targets (tests/golden/make_golden_fit2d.py)"""
	g = load_golden('ref2d_fit.npz')
	gv = fit2d(g, epochs, fused)
	assert gv.min_grid_scale == pytest.approx(float(g['min_grid_scale']), rel=1e-12) and gv.clamp_threshold == float(g['tau'])
	for nm in NAMES:
		want = g[f'after{epochs}_{nm}']
		got = getattr(gv, nm).detach().cpu().numpy().reshape(want.shape)
		d_ref, d_got = want - g[nm], got - g[nm].reshape(want.shape)
		assert np.abs(d_ref).max() > 0
		assert rel_err(got, want) < 1e-5, nm
		assert rel_err(d_got, d_ref) < 2e-2, (nm, rel_err(d_got, d_ref))
	assert gv.grid_scale == pytest.approx(float(g[f'after{epochs}_grid_scale']), rel=2e-6)
