"""
Pins the CPU oracle (oracle/) against golden vectors produced by executing the reference's own
kernel bodies and dense torch classes (tests/golden/make_golden.py).  CPU only.
"""
import numpy as np
import pytest

from helpers import NAMES, load_golden, oracle_from_golden, rel_err

TOL = {'f64': 1e-11, 'f32': 3e-5}


@pytest.mark.parametrize('prec', ['f64', 'f32'])
def test_grid3d(prec):
	g = load_golden(f'ref3d_kernels_{prec}.npz')
	o = oracle_from_golden(g, 3, prec)
	assert list(g['grid_size']) == o.dims
	assert np.float32(g['grid_scale']) == np.float32(o.grid_scale)
	np.testing.assert_array_equal(o.cnt, g['grid_cnt'])
	np.testing.assert_array_equal(o.offset, g['grid_offset'])
	np.testing.assert_array_equal(o.sorted_id, g['sorted_id'])


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('case', ['project', 'fit', 'boundary', 'all'])
def test_losses3d(prec, case):
	g = load_golden(f'ref3d_kernels_{prec}.npz')
	o = oracle_from_golden(g, 3, prec)
	val, grad = o.forward(g['in_x'])
	assert rel_err(val, g[f'{case}_val']) < TOL[prec]
	assert rel_err(grad, g[f'{case}_grad']) < TOL[prec]
	wv, wb, wg, wo, wh, wd = g[f'{case}_weights']
	separate = f'{case}_vor_positions' in g
	direct, vor, div = o.zero_grads(), (o.zero_grads() if separate else None), (o.zero_grads() if separate else None)
	o.backward3d(g['in_x'], val, grad, ref_val=g['in_ref_val'], weight_val=wv, normals=g['in_normals'], weight_boundary=wb,
				 ref_grad=g['in_ref_grad'], weight_grad=wg, ref_vor=g['in_ref_vor'], weight_vor=wo, ref_hel=g['in_ref_hel'], weight_hel=wh,
				 weight_div=wd, stop_gradient=g['in_stop_gradient'] if case == 'all' else None, direct=direct, vor=vor, div=div)
	for tag, grp in (('direct', direct), ('vor', vor), ('div', div)):
		if grp is None:
			continue
		for nm, a in zip(NAMES, grp):
			ref = g[f'{case}_{tag}_{nm}']
			if np.abs(ref).max() == 0:
				assert np.abs(a).max() == 0
			else:
				assert rel_err(a, ref) < TOL[prec], (tag, nm)


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('D', [2, 3])
def test_rk4_and_neighbors(prec, D):
	g = load_golden(f'ref{D}d_kernels_{prec}.npz')
	o = oracle_from_golden(g, D, prec)
	pos, deform, val, grad = o.rk4(g['in_x'], float(g['rk4_dt']), pos_only=False)
	for a, k in ((pos, 'rk4_pos'), (deform, 'rk4_deformation'), (val, 'rk4_val'), (grad, 'rk4_grad')):
		assert rel_err(a, g[k]) < TOL[prec], k
	assert rel_err(o.rk4(g['in_x'], float(g['rk4_dt'])), g['rk4_pos']) < TOL[prec]
	np.testing.assert_array_equal(o.mark_neighbors(g['in_x'][:3]), g['neighbors_mark'])


@pytest.mark.parametrize('prec', ['f64', 'f32'])
def test_grid2d(prec):
	g = load_golden(f'ref2d_kernels_{prec}.npz')
	o = oracle_from_golden(g, 2, prec)
	assert list(g['grid_size']) == o.dims
	np.testing.assert_array_equal(o.cnt, g['grid_cnt'])
	np.testing.assert_array_equal(o.offset, g['grid_offset'])
	np.testing.assert_array_equal(o.sorted_id, g['sorted_id'])


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('case', ['val', 'valb'])
def test_losses2d_value_kernel(prec, case):
	g = load_golden(f'ref2d_kernels_{prec}.npz')
	o = oracle_from_golden(g, 2, prec)
	val, _ = o.forward(g['in_x'], need_grad=False)
	assert rel_err(val, g[f'{case}_val']) < TOL[prec]
	w, wb = g[f'{case}_weights']
	direct = o.backward2d_val(g['in_x'], val, ref=g['in_ref'], weight=w, normals=g['in_normals'], normal_ref=g['in_normal_ref'], weight_boundary=wb,
							  stop_gradient=g['in_stop_gradient'] if case == 'valb' else None)
	for nm, a in zip(NAMES, direct):
		assert rel_err(a, g[f'{case}_direct_{nm}']) < TOL[prec], nm


@pytest.mark.parametrize('prec', ['f64', 'f32'])
@pytest.mark.parametrize('case', ['project', 'gall'])
def test_losses2d_gradient_kernel(prec, case):
	g = load_golden(f'ref2d_kernels_{prec}.npz')
	o = oracle_from_golden(g, 2, prec)
	_, grad = o.forward(g['in_x'], need_val=False)
	assert rel_err(grad, g[f'{case}_grad']) < TOL[prec]
	wg, wo, wd = g[f'{case}_weights']
	separate = f'{case}_vor_positions' in g
	direct, vor, div = o.zero_grads(), (o.zero_grads() if separate else None), (o.zero_grads() if separate else None)
	o.backward2d_grad(g['in_x'], grad, ref_grad=g['in_ref_grad'] if wg else None, weight_grad=wg, ref_vor=g['in_ref_vor'] if wo else None, weight_vor=wo,
					  weight_div=wd, stop_gradient=g['in_stop_gradient'] if case == 'gall' else None, direct=direct, vor=vor, div=div)
	for tag, grp in (('direct', direct), ('vor', vor), ('div', div)):
		if grp is None:
			continue
		for nm, a in zip(NAMES, grp):
			ref = g[f'{case}_{tag}_{nm}']
			if np.abs(ref).max() == 0:
				assert np.abs(a).max() == 0
			else:
				assert rel_err(a, ref) < TOL[prec], (tag, nm)


@pytest.mark.parametrize('D', [2, 3])
def test_dense_reference_class(D):
	"""tau = 0: the Fast path equals the reference's dense torch class (3D/GSR.py:118-130, 2D/GSR.py:115-147)."""
	from oracle.oracle import OracleGSR, extended_bounds
	g = load_golden('ref_dense_f64.npz')
	p = f'd{D}_in_'
	mgs = float(g[p + 'min_grid_scale'])
	ext = extended_bounds(D, (0., 1.) * D, mgs)
	o = OracleGSR(D, ext, g[p + 'positions'], g[p + 'scalings'], g[p + 'rotations'], g[p + 'values'], 0., mgs, precision='f64')
	assert o.cnt[0] == o.N	# one cell holds everything
	val, grad = o.forward(g[p + 'x'])
	assert rel_err(val, g[f'd{D}_val']) < 1e-11
	assert rel_err(grad, g[f'd{D}_grad']) < 1e-11


def test_init_fields_oracle_matches_reference_golden():
	"""N2: the oracle's regularised Biot-Savart sum against the reference's own vortex_particle / vortex_particle_gradient
	kernels run through the shim on the four 3D scenes (tests/golden/make_golden_init3d.py)"""
	import importlib
	import os
	import sys
	sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
	from helpers import ring_particles_np
	init_cond3d = importlib.import_module('gaussian_fluids_code_b200.init_cond3d')
	import oracle.oracle as orc
	g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref3d_init_fields.npz'))
	x = g['x']
	for name in ('leapfrog', 'single_vortex_ring', 'ring_collide', 'ring_with_obstacle'):
		for real, tag, tol in ((np.float64, 'f64', 1e-9), (np.float32, 'f32', 1e-4)):
			val, jac = np.zeros((x.shape[0], 3), real), np.zeros((x.shape[0], 3, 3), real)
			for ring in init_cond3d.rings_of(name):
				x0, w, U, a = ring_particles_np(ring, real)
				v, j = orc.vortex_particles(x, x0, w, U, a, real=real)
				val += v
				jac += j
			for got, key in ((val, 'val'), (jac, 'grad')):
				ref = g[f'{name}_{key}_{tag}']
				assert np.abs(got - ref).max() <= tol * np.abs(ref).max(), (name, key, tag, np.abs(got - ref).max() / np.abs(ref).max())


def test_mesh_sampler_oracle_matches_reference_golden():
	"""N3: the oracle's mesh sampler against the reference's own ti_get_tri_area / ti_lower_bound / ti_sample bodies run through
	the shim with a recorded sequence of uniforms (tests/golden/make_golden_mesh.py)"""
	import os
	import oracle.oracle as orc
	g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref3d_mesh_sampler.npz'))
	ps = orc.mesh_area_presum(g['vertices'], g['faces'])
	np.testing.assert_allclose(ps, g['area_presum'], rtol=2e-6)	# same serial order; the shim's cross product rounds each product separately too
	data, normal = orc.mesh_sample(g['uniforms'], g['vertices'], g['normals'], g['faces'], g['facenormals'], g['area_presum'])
	np.testing.assert_allclose(data, g['data'], rtol=0., atol=2e-7)
	np.testing.assert_allclose(normal, g['normal'], rtol=0., atol=5e-7)


def test_density_resampling_oracle_matches_reference_golden():
	"""N1: the oracle's trilinear resampling against the reference's own ti_get_interp_val body, at the back-traced lattice the
	reference's advection_rk4_ti produced (tests/golden/make_golden_density.py), and the whole composition of advected_density
	(RK4 back-trace by -dt, clamp, resample) through the oracle's RK4"""
	import os
	import oracle.oracle as orc
	from helpers import oracle_from_golden
	g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref3d_density.npz')))
	for real, tag, tol in ((np.float64, 'f64', 1e-12), (np.float32, 'f32', 2e-6)):
		for name in ('ring', 'smooth'):
			got = orc.interp_val(g[f'{name}_density_{tag}'], g[f'backtraced_{tag}'], g['domain'], real=real)
			assert np.abs(got - g[f'{name}_next_{tag}']).max() <= tol, (name, tag)
	# the composition, float64: lattice -> oracle RK4 by -dt -> clamp -> resample
	o = oracle_from_golden(g, 3, 'f64')
	o.build_grid()
	res, dom = g['res'], g['domain']
	axes = [np.linspace(dom[2 * a], dom[2 * a + 1], res[a], dtype=np.float32) for a in range(3)]	# get_grid_points: torch.linspace in float32
	x = np.stack(np.meshgrid(*axes, indexing='ij'), -1).reshape(-1, 3)
	bk = o.rk4(x, -float(g['dt']))
	bk = np.clip(bk, dom[0::2], dom[1::2]).reshape(*res, 3)
	assert np.abs(bk - g['backtraced_f64']).max() < 1e-6	# lattice coordinates: float32 linspace here, float64 in the f64 golden run
	got = orc.interp_val(g['smooth_density_f64'], bk, dom, real=np.float64)
	assert np.abs(got - g['smooth_next_f64']).max() < 2e-5


@pytest.mark.parametrize('epochs', [1, 3])
def test_optimisation_oracle_matches_reference_project(epochs):
	"""the oracle's restatement of project() + step() (OracleProjector3D: PCGrad, closed-form regulariser gradients, Adam,
	ReduceLROnPlateau, grid_scale) against the reference's OWN project() run through the shim (tests/golden/make_golden_project3d.py)"""
	import os
	import oracle.oracle as orc
	g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref3d_project.npz')))
	tau, mgs = float(g['tau']), float(g['min_grid_scale'])
	bounds = (0., 1.) * 3
	ext = orc.extended_bounds(3, bounds, mgs)
	prev = orc.OracleGSR(3, ext, g['cur_positions'], g['scalings'], g['rotations'], g['values'], tau, mgs, precision='f64')
	pr = orc.OracleProjector3D(bounds, [g['new_positions'], g['scalings'], g['rotations'], g['values']], prev, float(g['dt']), float(g['boundary_lambda']), tau, mgs)
	for k in range(epochs):
		pr.iterate(g['samples'][k], (g['boundary_data'][k], g['boundary_normal'][k]))
	before = dict(positions=g['new_positions'], scalings=g['scalings'], rotations=g['rotations'], values=g['values'])
	for nm, got in zip(('positions', 'scalings', 'rotations', 'values'), pr.params):
		want = g[f'after{epochs}_{nm}']
		d_ref, d_got = want.astype(np.float64) - before[nm], got - before[nm]
		assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), nm
		assert np.abs(d_got - d_ref).max() <= 2e-2 * np.abs(d_ref).max(), (nm, np.abs(d_got - d_ref).max() / np.abs(d_ref).max())
	assert pr.grid_scale == pytest.approx(float(g[f'after{epochs}_grid_scale']), rel=2e-6)


def test_optimisation_oracle_gradients_match_reference_project():
	"""the quantities the reference's step() consumes, iteration by iteration: the raw vorticity / divergence gradient sets as
	get_losses_ti leaves them, the total .grad after project()'s PCGrad projection + autograd regularisers + boundary pass, the
	scheduler metric and the learning rates in use — recorded inside the reference's own project() (make_golden_project3d.py).
	The parameter updates of the test above cannot see a wrong gradient magnitude (Adam's first step is lr * sign(g)); this can."""
	import os
	import oracle.oracle as orc
	g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref3d_project.npz')))
	tau, mgs = float(g['tau']), float(g['min_grid_scale'])
	bounds = (0., 1.) * 3
	prev = orc.OracleGSR(3, orc.extended_bounds(3, bounds, mgs), g['cur_positions'], g['scalings'], g['rotations'], g['values'], tau, mgs, precision='f64')
	pr = orc.OracleProjector3D(bounds, [g['new_positions'], g['scalings'], g['rotations'], g['values']], prev, float(g['dt']), float(g['boundary_lambda']), tau, mgs)
	for k in range(3):
		pr.iterate(g['samples'][k], (g['boundary_data'][k], g['boundary_normal'][k]))
		tol = 2e-5 * 3 ** k	# the golden is an f32 run; later iterations start from parameters that already differ by rounding
		for j, nm in enumerate(('positions', 'scalings', 'rotations', 'values')):
			for tag in ('vor', 'div'):
				want = g[f'it{k + 1}_{tag}_{nm}_grad']
				assert rel_err(pr.last[tag][j], want) < tol, (k, tag, nm, rel_err(pr.last[tag][j], want))
			want = g[f'it{k + 1}_total_{nm}_grad']
			assert rel_err(pr.last['total'][j], want) < tol, (k, 'total', nm, rel_err(pr.last['total'][j], want))
		assert pr.last['metric'] == pytest.approx(float(g[f'it{k + 1}_metric']), rel=1e-5)
		np.testing.assert_allclose(pr.last['lr'], g[f'it{k + 1}_lr_used'], rtol=1e-12)


@pytest.mark.parametrize('epochs', [1, 3])
def test_optimisation_oracle_matches_reference_project_2d(epochs):
	"""OracleProjector2D against the reference's OWN 2D project() run through the shim (tests/golden/make_golden_project2d.py)"""
	import os
	import oracle.oracle as orc
	g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref2d_project.npz')))
	tau, mgs = float(g['tau']), float(g['min_grid_scale'])
	dom = tuple(float(v) for v in g['domain'])
	ext = orc.extended_bounds(2, dom, mgs)
	prev = orc.OracleGSR(2, ext, g['cur_positions'], g['scalings'], g['rotations'], g['values'], tau, mgs, precision='f64')
	pr = orc.OracleProjector2D(dom, [g['new_positions'], g['scalings'], g['rotations'], g['values']], prev, float(g['dt']), float(g['boundary_lambda']), tau, mgs, dom)
	for k in range(epochs):
		pr.iterate(g['samples'][k], (g['b1_data'][k], g['b1_val'][k]), (g['b2_data'][k], g['b2_normal'][k], g['b2_ref'][k]))
	before = dict(positions=g['new_positions'], scalings=g['scalings'], rotations=g['rotations'], values=g['values'])
	for nm, got in zip(('positions', 'scalings', 'rotations', 'values'), pr.params):
		want = g[f'after{epochs}_{nm}']
		got = got.reshape(want.shape)
		d_ref, d_got = want.astype(np.float64) - before[nm], got - before[nm]
		assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), nm
		assert np.abs(d_got - d_ref).max() <= 2e-2 * np.abs(d_ref).max(), (nm, np.abs(d_got - d_ref).max() / np.abs(d_ref).max())
	assert pr.grid_scale == pytest.approx(float(g[f'after{epochs}_grid_scale']), rel=2e-6)


@pytest.mark.parametrize('epochs', [1, 3])
def test_fit_oracle_matches_reference_fit(epochs):
	"""OracleFit3D against the reference's OWN fit_velocity_with_gradient run through the shim (tests/golden/make_golden_fit3d.py)"""
	import os
	import oracle.oracle as orc
	g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref3d_fit.npz')))
	bounds = (0., 1.) * 3
	N = g['positions'].shape[0]
	mgs = orc.default_min_grid_scale(3, bounds, N)
	fit = orc.OracleFit3D(bounds, [g['positions'], g['scalings'], g['rotations'], g['values']], g['lrs'], 5e-3, mgs)
	for k in range(epochs):
		fit.iterate(g['samples'][k], g['ref_val'][k], g['ref_grad'][k])
	for nm, got in zip(('positions', 'scalings', 'rotations', 'values'), fit.params):
		want = g[f'after{epochs}_{nm}']
		d_ref, d_got = want.astype(np.float64) - g[nm], got - g[nm]
		assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), nm
		assert np.abs(d_got - d_ref).max() <= 2e-2 * np.abs(d_ref).max(), (nm, np.abs(d_got - d_ref).max() / np.abs(d_ref).max())
	assert fit.grid_scale == pytest.approx(float(g[f'after{epochs}_grid_scale']), rel=2e-6)


def test_oracle_adam_and_plateau_scheduler_match_torch():
	"""the oracle's Adam / ReduceLROnPlateau restatement (oracle._Adam) against torch.optim on a random gradient and metric
	sequence long enough for the scheduler to fire several times (the project goldens are too short for that)"""
	import torch
	import oracle.oracle as orc
	rng = np.random.default_rng(11)
	p0 = rng.normal(size=(7, 3))
	t = torch.tensor(p0, dtype=torch.float64, requires_grad=True)
	opt = torch.optim.Adam([t], lr=3e-3)
	sch = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=.9, patience=4)
	mine = orc._Adam(3e-3, patience=4, factor=.9)
	p = p0.copy()
	metric = 1.
	fired = 0
	for k in range(200):
		g = rng.normal(size=p0.shape) * (1. if k % 7 else 1e-9)	# now and then a vanishing gradient: the eps term matters
		metric = metric * (.97 if (k // 20) % 2 == 0 else 1.01)	# improving, then stagnating, in turns
		t.grad = torch.tensor(g)
		opt.step()
		sch.step(metric)
		p = mine.step(p, g)
		lr_before = mine.lr
		mine.schedule(metric)
		fired += mine.lr != lr_before
		assert mine.lr == pytest.approx(opt.param_groups[0]['lr'], rel=1e-12), k
		np.testing.assert_allclose(p, t.detach().numpy(), rtol=1e-10, atol=1e-14)
	assert fired >= 3


@pytest.mark.parametrize('epochs', [1, 3])
def test_fit_oracle_matches_reference_fit_2d(epochs):
	"""OracleFit2D against the reference's OWN 2D fit_velocity_with_gradient run through the shim (tests/golden/make_golden_fit2d.py)"""
	import os
	import oracle.oracle as orc
	g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref2d_fit.npz')))
	dom = tuple(float(v) for v in g['domain'])
	fit = orc.OracleFit2D(dom, [g['positions'], g['scalings'], g['rotations'], g['values']], g['lrs'], float(g['tau']), float(g['min_grid_scale']))
	for k in range(epochs):
		fit.iterate(g['samples'][k], g['ref_val'][k], g['ref_grad'][k])
	for nm, got in zip(('positions', 'scalings', 'rotations', 'values'), fit.params):
		want = g[f'after{epochs}_{nm}']
		got = got.reshape(want.shape)
		d_ref, d_got = want.astype(np.float64) - g[nm], got - g[nm]
		assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), nm
		assert np.abs(d_got - d_ref).max() <= 2e-2 * np.abs(d_ref).max(), (nm, np.abs(d_got - d_ref).max() / np.abs(d_ref).max())
	assert fit.grid_scale == pytest.approx(float(g[f'after{epochs}_grid_scale']), rel=2e-6)


def test_dense_torch_restatement_matches_reference_dense_class():
	"""oracle.dense_torch_value_gradient (the CPU baseline B2 of bench.py) against the golden made by executing the reference's
	own GaussianSplatting3D.__call__ / gradient (3D/GSR.py:93-130) in float64"""
	import torch
	from oracle.oracle import dense_torch_value_gradient
	g = load_golden('ref_dense_f64.npz')
	T = lambda k: torch.tensor(g[k], dtype=torch.float64)
	u, grad = dense_torch_value_gradient(T('d3_in_positions'), T('d3_in_scalings'), T('d3_in_rotations'), T('d3_in_values'), T('d3_in_x'))
	assert rel_err(u.numpy(), g['d3_val']) < 1e-12 and rel_err(grad.numpy(), g['d3_grad']) < 1e-12


def _check_steps(g, last, k, sets=True):
	"""one iteration's pre-step quantities against the golden recorded inside the reference's own loop (see the 3D project test)"""
	tol = 2e-5 * 3 ** k
	for j, nm in enumerate(('positions', 'scalings', 'rotations', 'values')):
		for tag in (('vor', 'div') if sets else ()):
			want = g[f'it{k + 1}_{tag}_{nm}_grad']
			assert rel_err(last[tag][j].reshape(want.shape), want) < tol, (k, tag, nm, rel_err(last[tag][j].reshape(want.shape), want))
		want = g[f'it{k + 1}_total_{nm}_grad']
		assert rel_err(last['total'][j].reshape(want.shape), want) < tol, (k, 'total', nm, rel_err(last['total'][j].reshape(want.shape), want))
	assert last['metric'] == pytest.approx(float(g[f'it{k + 1}_metric']), rel=1e-5)
	np.testing.assert_allclose(last['lr'], g[f'it{k + 1}_lr_used'], rtol=1e-12)


def test_optimisation_oracle_gradients_match_reference_project_2d():
	"""2D: raw vorticity / divergence sets of get_grad_losses, total .grad (value samples, PCGrad, normal samples, regularisers
	incl. the position drift), metric and lrs at every step() of the reference's own 2D project()"""
	import oracle.oracle as orc
	g = load_golden('ref2d_project.npz')
	tau, mgs = float(g['tau']), float(g['min_grid_scale'])
	dom = tuple(float(v) for v in g['domain'])
	prev = orc.OracleGSR(2, orc.extended_bounds(2, dom, mgs), g['cur_positions'], g['scalings'], g['rotations'], g['values'], tau, mgs, precision='f64')
	pr = orc.OracleProjector2D(dom, [g['new_positions'], g['scalings'], g['rotations'], g['values']], prev, float(g['dt']), float(g['boundary_lambda']), tau, mgs, dom)
	for k in range(3):
		pr.iterate(g['samples'][k], (g['b1_data'][k], g['b1_val'][k]), (g['b2_data'][k], g['b2_normal'][k], g['b2_ref'][k]))
		_check_steps(g, pr.last, k)


@pytest.mark.parametrize('D', [3, 2])
def test_fit_oracle_gradients_match_reference_fit(D):
	"""the total .grad, the metric and the lrs at every step() of the reference's own fit_velocity_with_gradient (3D and 2D)"""
	import oracle.oracle as orc
	g = load_golden(f'ref{D}d_fit.npz')
	if D == 3:
		bounds = (0., 1.) * 3
		fit = orc.OracleFit3D(bounds, [g['positions'], g['scalings'], g['rotations'], g['values']], g['lrs'], 5e-3, orc.default_min_grid_scale(3, bounds, g['positions'].shape[0]))
	else:
		fit = orc.OracleFit2D(tuple(float(v) for v in g['domain']), [g['positions'], g['scalings'], g['rotations'], g['values']], g['lrs'], float(g['tau']), float(g['min_grid_scale']))
	for k in range(3):
		fit.iterate(g['samples'][k], g['ref_val'][k], g['ref_grad'][k])
		_check_steps(g, fit.last, k, sets=False)


def test_split_restatement_matches_reference_clone_2d():
	"""oracle.split_gaussians (the reseeding split of clone_velocity_field) against the golden recorded inside the reference's own
	2D clone_velocity_field (tests/golden/make_golden_clone2d.py): the field right after the split for the recorded normal draws"""
	import oracle.oracle as orc
	g = load_golden('ref2d_clone.npz')
	P, S, R, V, stop = orc.split_gaussians(2, g['positions'], g['scalings'], g['rotations'], g['values'], g['normals'])
	assert P.shape[0] == g['split_positions'].shape[0] == g['positions'].shape[0] + g['normals'].shape[1]
	assert rel_err(P, g['split_positions']) < 2e-6 and rel_err(S, g['split_scalings']) < 1e-6
	assert rel_err(R.reshape(-1), g['split_rotations'].reshape(-1)) < 1e-7 and rel_err(V, g['split_values']) < 1e-7
	# untouched Gaussians far from every child stay frozen in the golden; children never are
	assert not g['stop_gradient'][~stop].any() and (g['stop_gradient'] <= stop).all()


def test_clone_refit_gradients_match_reference_clone_2d():
	"""the first refit iteration of the reference's 2D clone_velocity_field (2D/advance.py:96-120): value + gradient L1 losses of
	the split field against the source field with the frozen Gaussians skipped (GSR.py:291, :404).  The recorded batches lie inside
	the split parents' support — elsewhere res == src up to rounding and sign(val - ref) is a coin flip in any implementation."""
	import oracle.oracle as orc
	g = load_golden('ref2d_clone.npz')
	tau, mgs = float(g['tau']), float(g['min_grid_scale'])
	ext = orc.extended_bounds(2, tuple(float(v) for v in g['domain']), mgs)
	src = orc.OracleGSR(2, ext, g['positions'], g['scalings'], g['rotations'], g['values'], tau, mgs)
	res = orc.OracleGSR(2, ext, g['split_positions'], g['split_scalings'], g['split_rotations'], g['split_values'], tau, mgs)
	assert np.array_equal(res.mark_neighbors(g['split_positions'][-2 * g['normals'].shape[1]:]).astype(bool) | (np.arange(res.N) >= g['positions'].shape[0] - g['normals'].shape[1]),
						  ~g['stop_gradient'].astype(bool))
	x = g['samples'][0]
	ref_val, ref_grad = src.forward(x)
	val, grad = res.forward(x)
	d = res.zero_grads()
	sg = g['stop_gradient'].astype(np.int32)
	res.backward2d_val(x, val, ref=ref_val, weight=1., stop_gradient=sg, direct=d)
	res.backward2d_grad(x, grad, ref_grad=ref_grad, weight_grad=1., stop_gradient=sg, direct=d)
	for k, nm in enumerate(NAMES):
		if nm == 'scalings':	# + the autograd regularisers, covered by the CUDA test
			continue
		want = g[f'it1_total_{nm}_grad']
		assert np.abs(want).max() > 0
		assert rel_err(np.asarray(d[k]).reshape(want.shape), want) < 2e-5, nm
