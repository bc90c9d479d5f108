"""
TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by executing the REFERENCE'S OWN code:

  * the Taichi kernel bodies of /root/reference/3D/GSR.py and /root/reference/2D/GSR.py, run as
    plain Python through tests/golden/ti_shim.py (taichi itself is not installed), once in
    float32 (the reference's arithmetic) and once in float64 (pins the formulas to ~1e-13);
  * the reference's dense torch classes GaussianSplatting3D / GaussianSplatting (tau = 0).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The resulting fixtures are committed; tests/test_oracle_golden.py checks oracle/ against them and
the GPU parity tests check the CUDA path against them.  Nothing here is copied from the reference:
the script imports it.
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ti_shim  # noqa: E402
from ti_shim import GArr  # noqa: E402

REF = '/root/reference'


def load_ref(sub, name):
	ti_shim.install()
	argv = sys.argv
	sys.argv = ['x', '--device', 'cpu', '--dir', tempfile.mkdtemp()]
	try:
		spec = importlib.util.spec_from_file_location(name, os.path.join(REF, sub, 'GSR.py'))
		mod = importlib.util.module_from_spec(spec)
		spec.loader.exec_module(mod)
	finally:
		sys.argv = argv
	return mod


def f32r(a):
	"""values exactly representable in float32, stored as float64"""
	return np.asarray(a, np.float32).astype(np.float64)


def scene(D, rng, n_per_axis, Q, tau, aniso=.15):
	"""jittered lattice of Gaussians in [0,1]^D with random shapes; Q random samples (some outside the domain)"""
	axes = [np.linspace(0., 1., n) for n in n_per_axis]
	P = np.stack(np.meshgrid(*axes, indexing='ij'), -1).reshape(-1, D)
	h = 1. / (max(n_per_axis) - 1)
	P = P + rng.uniform(-.25 * h, .25 * h, P.shape)
	N = P.shape[0]
	mgs = .5 if D == 3 else .75
	s0 = .5 * np.log(-2. * np.log(tau)) - np.log(mgs) if tau else -np.log(mgs) + .5
	S = s0 + rng.uniform(-aniso, aniso, (N, D))
	R = rng.normal(size=(N, 4)) if D == 3 else rng.uniform(-np.pi, np.pi, N)
	V = rng.normal(scale=.3, size=(N, D))
	X = rng.uniform(-.05, 1.05, (Q, D))
	return dict(positions=f32r(P), scalings=f32r(S), rotations=f32r(R), values=f32r(V), x=f32r(X), min_grid_scale=mgs, tau=float(np.float32(tau)))


def new_fast(mod, D, sc, dt):
	cls = mod.GaussianSplatting3DFast if D == 3 else mod.GaussianSplattingFast
	o = object.__new__(cls)
	o.N, o.dim = sc['positions'].shape[0], D
	o.min_grid_scale, o.clamp_threshold = sc['min_grid_scale'], sc['tau']
	m = o.min_grid_scale
	o.x_min, o.x_max, o.y_min, o.y_max = 0. - m, 1. + m, 0. - m, 1. + m
	if D == 3:
		o.z_min, o.z_max = 0. - m, 1. + m
	o.create_grid_data()
	# reinitialize_grid's host formula (3D/GSR.py:247-251) on the f32 scalings
	if o.clamp_threshold:
		o.grid_scale = max(np.sqrt(-2. * np.log(o.clamp_threshold)) * np.exp(-float(np.float32(sc['scalings']).min())), o.min_grid_scale)
	else:
		o.grid_scale = max(o.x_max - o.x_min, o.y_max - o.y_min, *( [o.z_max - o.z_min] if D == 3 else [] ))
	P = [GArr(sc[k], dtype=dt) for k in ('positions', 'scalings', 'rotations', 'values')]
	o.reinitialize_grid_ti(P[0], o.grid_scale)
	return o, P


def grid_out(o):
	return dict(grid_cnt=o.grid_cnt.arr.reshape(-1).copy(), grid_offset=o.grid_offset.arr.reshape(-1).copy(),
				sorted_id=o.sorted_id.arr[:o.N].copy(), grid_size=np.array(o.grid_size), grid_scale=o.grid_scale)


def run3d(mod, dt, seed):
	ti_shim.set_dtype(dt)
	rng = np.random.default_rng(seed)
	out = {}
	sc = scene(3, rng, (4, 4, 3), 14, 5e-3)
	out.update({'in_' + k: v for k, v in sc.items()})
	o, P = new_fast(mod, 3, sc, dt)
	out.update(grid_out(o))
	Q, N = sc['x'].shape[0], o.N
	x = np.array(sc['x'], dt)
	ins = dict(ref_val=f32r(rng.normal(scale=.2, size=(Q, 3))), normals=f32r(rng.normal(size=(Q, 3))),
			   ref_grad=f32r(rng.normal(scale=.5, size=(Q, 3, 3))), ref_vor=f32r(rng.normal(scale=.5, size=(Q, 3))),
			   ref_hel=f32r(rng.normal(scale=.1, size=Q)), stop_gradient=(rng.uniform(size=N) < .15).astype(np.int32))
	out.update({'in_' + k: v for k, v in ins.items()})
	Z = lambda *s: np.zeros(s, dt)

	def call(weights, sets, stop=None):
		"""weights = (val, boundary, grad, vor, hel, div); sets = 'separate' | 'alias'"""
		for p in P:
			p.grad = Z(*p.shape)
		direct = [p.grad for p in P]
		vor = [Z(*p.shape) for p in P] if sets == 'separate' else direct
		div = [Z(*p.shape) for p in P] if sets == 'separate' else direct
		val, grad = Z(Q, 3), Z(Q, 3, 3)
		wv, wb, wg, wo, wh, wd = weights
		o.get_losses_ti(P[0], P[1], P[2], P[3], o.grid_scale, x,
						np.array(ins['ref_val'], dt) if wv else Z(Q, 3), wv,
						np.array(ins['normals'], dt) if wb else Z(Q, 3), wb,
						np.array(ins['ref_grad'], dt) if wg else Z(Q, 3, 3), wg,
						np.array(ins['ref_vor'], dt) if wo else Z(Q, 3), wo,
						np.array(ins['ref_hel'], dt) if wh else Z(Q), wh, wd,
						val, grad, *vor, *div,
						stop if stop is not None else np.zeros(N, np.int32))
		return val, grad, direct, vor, div

	cases = {'project': ((0., 0., 0., 1., 1., 1.), 'separate', None),
			 'fit': ((1., 0., 1., 0., 0., 0.), 'alias', None),
			 'boundary': ((0., 10., 0., 0., 0., 0.), 'alias', None),
			 'all': ((.7, 3., 1.3, .9, 1.1, .6), 'alias', ins['stop_gradient'])}
	for name, (w, sets, stop) in cases.items():
		val, grad, direct, vor, div = call(w, sets, stop)
		out[f'{name}_weights'] = np.array(w)
		out[f'{name}_val'], out[f'{name}_grad'] = val, grad
		for tag, grp in (('direct', direct), ('vor', vor), ('div', div)):
			if tag != 'direct' and sets == 'alias':
				continue
			for nm, a in zip(('positions', 'scalings', 'rotations', 'values'), grp):
				out[f'{name}_{tag}_{nm}'] = a.copy()
	# RK4 (full outputs) and pos-only
	dtv = -.37
	goal, deform, gval, ggrad = Z(Q, 3), Z(Q, 3, 3), Z(Q, 3), Z(Q, 3, 3)
	o.advection_rk4_ti(P[0], P[1], P[2], P[3], o.grid_scale, x, dtv, goal, deform, gval, ggrad)
	out.update(rk4_dt=dtv, rk4_pos=goal, rk4_deformation=deform, rk4_val=gval, rk4_grad=ggrad)
	mark = np.zeros(N, np.int32)
	o.get_all_neighbors_ti(x[:3], P[0], o.grid_scale, mark)
	out['neighbors_mark'] = mark
	return out


def run2d(mod, dt, seed):
	ti_shim.set_dtype(dt)
	rng = np.random.default_rng(seed)
	out = {}
	sc = scene(2, rng, (5, 4), 14, 1e-3)
	out.update({'in_' + k: v for k, v in sc.items()})
	o, P = new_fast(mod, 2, sc, dt)
	out.update(grid_out(o))
	Q, N = sc['x'].shape[0], o.N
	x = np.array(sc['x'], dt)
	ins = dict(ref=f32r(rng.normal(scale=.2, size=(Q, 2))), normals=f32r(rng.normal(size=(Q, 2))), normal_ref=f32r(rng.normal(scale=.1, size=Q)),
			   ref_grad=f32r(rng.normal(scale=.5, size=(Q, 2, 2))), ref_vor=f32r(rng.normal(scale=.5, size=Q)),
			   stop_gradient=(rng.uniform(size=N) < .15).astype(np.int32))
	out.update({'in_' + k: v for k, v in ins.items()})
	Z = lambda *s: np.zeros(s, dt)
	names = ('positions', 'scalings', 'rotations', 'values')
	# value kernel: value-L1 + boundary
	for name, (w, wb, stop) in {'val': ((1., 0., None)), 'valb': ((.8, 2.5, ins['stop_gradient']))}.items():
		for p in P:
			p.grad = Z(*p.shape)
		val = np.full((Q, 2), 7., dt)	# the kernel zeroes it
		o.get_losses_ti(P[0], P[1], P[2], P[3], o.grid_scale, x, np.array(ins['ref'], dt), w,
						np.array(ins['normals'], dt), np.array(ins['normal_ref'], dt), wb, val,
						stop if stop is not None else np.zeros(N, np.int32))
		out[f'{name}_weights'] = np.array([w, wb])
		out[f'{name}_val'] = val
		for nm, p in zip(names, P):
			out[f'{name}_direct_{nm}'] = p.grad.copy()
	# gradient kernel
	for name, (w, sets, stop) in {'project': ((0., 1., 1.), 'separate', None), 'gall': ((1.2, .7, .9), 'alias', ins['stop_gradient'])}.items():
		for p in P:
			p.grad = Z(*p.shape)
		direct = [p.grad for p in P]
		vor = [Z(*p.shape) for p in P] if sets == 'separate' else direct
		div = [Z(*p.shape) for p in P] if sets == 'separate' else direct
		grad = np.full((Q, 2, 2), 7., dt)
		o.get_grad_losses_ti(P[0], P[1], P[2], P[3], o.grid_scale, x, np.array(ins['ref_grad'], dt) if w[0] else Z(Q, 2, 2), w[0],
							 np.array(ins['ref_vor'], dt) if w[1] else Z(Q), w[1], w[2], grad, *vor, *div,
							 stop if stop is not None else np.zeros(N, np.int32))
		out[f'{name}_weights'] = np.array(w)
		out[f'{name}_grad'] = grad
		for tag, grp in (('direct', direct), ('vor', vor), ('div', div)):
			if tag != 'direct' and sets == 'alias':
				continue
			for nm, a in zip(names, grp):
				out[f'{name}_{tag}_{nm}'] = a.copy()
	dtv = .41
	goal, deform, gval, ggrad = Z(Q, 2), Z(Q, 2, 2), Z(Q, 2), Z(Q, 2, 2)
	o.advection_rk4_ti(P[0], P[1], P[2], P[3], o.grid_scale, x, dtv, goal, deform, gval, ggrad)
	out.update(rk4_dt=dtv, rk4_pos=goal, rk4_deformation=deform, rk4_val=gval, rk4_grad=ggrad)
	mark = np.zeros(N, np.int32)
	o.get_all_neighbors_ti(x[:3], P[0], o.grid_scale, mark)
	out['neighbors_mark'] = mark
	return out


def run_dense(mod3, mod2, seed):
	"""the reference's dense torch classes (untruncated sum == the Fast path with tau = 0)"""
	import torch
	rng = np.random.default_rng(seed)
	out = {}
	for D, mod in ((3, mod3), (2, mod2)):
		sc = scene(D, rng, (4, 4, 3) if D == 3 else (5, 4), 12, 0.)
		cls = mod.GaussianSplatting3D if D == 3 else mod.GaussianSplatting
		o = object.__new__(cls)
		o.N, o.dim = sc['positions'].shape[0], D
		for k in ('positions', 'scalings', 'rotations', 'values'):
			setattr(o, k, torch.tensor(sc[k], dtype=torch.float64))
		mod.device = torch.device('cpu')
		with torch.no_grad():
			# the classes allocate R with the default dtype; run them in float64 by switching it
			torch.set_default_dtype(torch.float64)
			try:
				grad, val = o.gradient(torch.tensor(sc['x'], dtype=torch.float64), need_val=True)
				cov_inv = o.get_variances()
			finally:
				torch.set_default_dtype(torch.float32)
		out.update({f'd{D}_in_{k}': v for k, v in sc.items()})
		out[f'd{D}_val'], out[f'd{D}_grad'], out[f'd{D}_cov_inv'] = val.numpy(), grad.numpy(), cov_inv.numpy()
	return out


if __name__ == '__main__':
	mod3 = load_ref('3D', 'ref3d_GSR')
	mod2 = load_ref('2D', 'ref2d_GSR')
	for tag, dt in (('f64', np.float64), ('f32', np.float32)):
		np.savez_compressed(os.path.join(HERE, f'ref3d_kernels_{tag}.npz'), **run3d(mod3, dt, 1234))
		np.savez_compressed(os.path.join(HERE, f'ref2d_kernels_{tag}.npz'), **run2d(mod2, dt, 4321))
		print('wrote', tag)
	np.savez_compressed(os.path.join(HERE, 'ref_dense_f64.npz'), **run_dense(mod3, mod2, 99))
	print('done')
