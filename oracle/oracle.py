"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes binding of the CPU oracle (oracle/gsr_oracle.c), a C restatement of the reference's
Taichi kernels (3D/GSR.py, 2D/GSR.py), plus numpy restatements of the reference's small
host-side formulas.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module; the product package never does.

Parity status: pinned — see tests/golden/make_golden.py (golden vectors produced by running the
reference's own kernel bodies) and tests/test_oracle_golden.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libgsr_oracle.so')
_lib = None


def build(force=False):
	"""Compile the oracle with the system gcc (`make -C oracle`)."""
	srcs = [os.path.join(_HERE, f) for f in ('gsr_oracle.c', 'gsr3d_oracle_impl.h', 'gsr2d_oracle_impl.h', 'Makefile')]
	if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
		subprocess.run(['make', '-C', _HERE], check=True, stdout=subprocess.DEVNULL)
	return _SO


def lib():
	global _lib
	if _lib is None:
		build()
		_lib = C.CDLL(_SO)
	return _lib


class _Grid3(C.Structure):
	_fields_ = [('lo', C.c_float * 3), ('hi', C.c_float * 3), ('grid_scale', C.c_float), ('dims', C.c_int * 3),
				('cnt', C.c_void_p), ('offset', C.c_void_p), ('sorted_id', C.c_void_p)]


class _Grid2(C.Structure):
	_fields_ = [('lo', C.c_float * 2), ('hi', C.c_float * 2), ('grid_scale', C.c_float), ('dims', C.c_int * 2),
				('cnt', C.c_void_p), ('offset', C.c_void_p), ('sorted_id', C.c_void_p)]


def _p(a):
	return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
	return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------------------------------------
# host-side formulas of the reference
# ---------------------------------------------------------------------------------------------

def default_min_grid_scale(D, bounds, N):
	"""3D/GSR.py:160 (2 * (V/N)^(1/3)); 2D/GSR.py:177 (3 * (A/N)^(1/2)). bounds = (x_min, x_max, y_min, ...)."""
	ext = [bounds[2 * k + 1] - bounds[2 * k] for k in range(D)]
	if D == 3:
		return (ext[0] * ext[1] * ext[2] / N) ** (1. / 3.) * 2.
	return (ext[0] * ext[1] / N) ** .5 * 3.


def extended_bounds(D, bounds, min_grid_scale):
	"""3D/GSR.py:162-164; 2D/GSR.py:179."""
	out = []
	for k in range(D):
		out += [bounds[2 * k] - min_grid_scale, bounds[2 * k + 1] + min_grid_scale]
	return tuple(out)


def initial_scaling(tau, min_grid_scale):
	"""3D/GSR.py:166; 2D/GSR.py:181."""
	return .5 * np.log(-2. * np.log(tau)) - np.log(min_grid_scale)


def grid_dims(D, ext_bounds, min_grid_scale):
	"""create_grid_data: 3D/GSR.py:173 (all axes `//`), 2D/GSR.py:188 (x uses `//`, y uses `/`)."""
	if D == 3:
		return [int((ext_bounds[2 * k + 1] - ext_bounds[2 * k]) // min_grid_scale) + 1 for k in range(3)]
	return [int((ext_bounds[1] - ext_bounds[0]) // min_grid_scale) + 1, int((ext_bounds[3] - ext_bounds[2]) / min_grid_scale) + 1]


def grid_scale_of(tau, scalings, min_grid_scale, ext_bounds):
	"""reinitialize_grid: 3D/GSR.py:247-251; 2D/GSR.py:224-228 (host double arithmetic on the f32 min)."""
	if tau:
		return max(np.sqrt(-2. * np.log(tau)) * np.exp(-float(np.min(scalings))), min_grid_scale)
	D = len(ext_bounds) // 2
	return max(ext_bounds[2 * k + 1] - ext_bounds[2 * k] for k in range(D))


# ---------------------------------------------------------------------------------------------
# field objects
# ---------------------------------------------------------------------------------------------

class OracleGSR:
	"""CPU oracle of GaussianSplatting3DFast (D=3) / GaussianSplattingFast (D=2) for fixed parameters."""

	def __init__(self, D, ext_bounds, positions, scalings, rotations, values, tau, min_grid_scale, dims=None, grid_scale=None, precision='f32', nthreads=1):
		assert D in (2, 3)
		self.D = D
		self.sfx = '_' + precision
		self.real = np.float32 if precision == 'f32' else np.float64
		self.creal = C.c_float if precision == 'f32' else C.c_double
		self.nthreads = int(nthreads)
		self.ext_bounds = tuple(float(b) for b in ext_bounds)
		self.positions, self.scalings, self.rotations, self.values = _f32(positions), _f32(scalings), _f32(rotations), _f32(values)
		self.N = self.positions.shape[0]
		self.dim = self.values.shape[1]
		self.tau = float(tau)
		self.min_grid_scale = float(min_grid_scale)
		self.dims = list(dims) if dims is not None else grid_dims(D, self.ext_bounds, self.min_grid_scale)
		self.grid_scale = float(grid_scale) if grid_scale is not None else grid_scale_of(self.tau, self.scalings, self.min_grid_scale, self.ext_bounds)
		self.build_grid()

	def build_grid(self):
		D = self.D
		ncell = int(np.prod(self.dims))
		self.cnt = np.zeros(ncell, np.int32)
		self.offset = np.zeros(ncell, np.int32)
		self.sorted_id = np.full(max(self.N, 1), -1, np.int32)
		lo = np.array([self.ext_bounds[2 * k] for k in range(D)], np.float32)
		hi = np.array([self.ext_bounds[2 * k + 1] for k in range(D)], np.float32)
		dims = np.array(self.dims, np.int32)
		fn = lib().o3_build_grid if D == 3 else lib().o2_build_grid
		fn.restype = C.c_int
		self.n_in = fn(_p(self.positions), C.c_long(self.N), _p(lo), _p(hi), C.c_float(np.float32(self.grid_scale)), _p(dims),
					   _p(self.cnt), _p(self.offset), _p(self.sorted_id))
		G = _Grid3() if D == 3 else _Grid2()
		for k in range(D):
			G.lo[k], G.hi[k], G.dims[k] = lo[k], hi[k], int(dims[k])
		G.grid_scale = np.float32(self.grid_scale)
		G.cnt, G.offset, G.sorted_id = self.cnt.ctypes.data, self.offset.ctypes.data, self.sorted_id.ctypes.data
		self._G = G
		return self.cnt, self.offset, self.sorted_id

	def _fn(self, name):
		f = getattr(lib(), f'o{self.D}_{name}{self.sfx}')
		f.restype = None
		return f

	def forward(self, x, need_grad=True, need_val=True):
		"""loop 1 of get_losses_ti — returns (val (Q,dim), grad (Q,dim,D) | None)."""
		x = _f32(x)
		Q, D, dim = x.shape[0], self.D, self.dim
		val = np.zeros((Q, dim), self.real)
		grad = np.zeros((Q, dim, D), self.real) if need_grad else None
		self._fn('forward')(C.byref(self._G), _p(self.positions), _p(self.scalings), _p(self.rotations), _p(self.values),
							C.c_double(self.tau), C.c_int(dim), _p(x), C.c_long(Q), _p(val), _p(grad), C.c_int(self.nthreads))
		return val, grad

	def zero_grads(self):
		return [np.zeros(a.shape, self.real) for a in (self.positions, self.scalings, self.rotations, self.values)]

	def backward3d(self, x, val, grad, ref_val=None, weight_val=0., normals=None, weight_boundary=0., ref_grad=None, weight_grad=0.,
				   ref_vor=None, weight_vor=0., ref_hel=None, weight_hel=0., weight_div=0., stop_gradient=None,
				   direct=None, vor=None, div=None):
		"""loop 2 of 3D get_losses_ti.  direct/vor/div are lists of 4 arrays (pos, scal, rot, val grads) accumulated into;
		vor/div default to `direct` (the aliasing of 3D/GSR.py:564-579)."""
		assert self.D == 3
		x = _f32(x)
		Q, dim = x.shape[0], self.dim
		z = lambda *s: np.zeros(s, np.float32)
		ref_val = _f32(ref_val) if weight_val != 0. else z(Q, dim)
		normals = _f32(normals) if weight_boundary != 0. else z(Q, dim)
		ref_grad = _f32(ref_grad) if weight_grad != 0. else z(Q, dim, 3)
		ref_vor = _f32(ref_vor) if weight_vor != 0. else z(Q, 3)
		ref_hel = _f32(ref_hel) if weight_hel != 0. else z(Q)
		direct = direct if direct is not None else self.zero_grads()
		vor = vor if vor is not None else direct
		div = div if div is not None else direct
		w = np.array([weight_val, weight_boundary, weight_grad, weight_vor, weight_hel, weight_div], np.float64)
		sg = None if stop_gradient is None else np.ascontiguousarray(stop_gradient, dtype=np.int32)
		val = np.ascontiguousarray(val, dtype=self.real)
		grad = np.ascontiguousarray(grad, dtype=self.real)
		self._fn('backward')(C.byref(self._G), _p(self.positions), _p(self.scalings), _p(self.rotations), _p(self.values),
							 C.c_double(self.tau), C.c_int(dim), _p(x), C.c_long(Q), _p(val), _p(grad),
							 _p(ref_val), _p(normals), _p(ref_grad), _p(ref_vor), _p(ref_hel), _p(w), _p(sg),
							 *[_p(a) for a in direct], *[_p(a) for a in vor], *[_p(a) for a in div], C.c_int(self.nthreads))
		return direct, vor, div

	def backward2d_val(self, x, val, ref=None, weight=0., normals=None, normal_ref=None, weight_boundary=0., stop_gradient=None, direct=None):
		"""loop 2 of the 2D value kernel (2D/GSR.py:282-339) with the wrapper's defaults (:341-350)."""
		assert self.D == 2
		x = _f32(x)
		Q, dim = x.shape[0], self.dim
		if ref is None:
			ref, weight = np.zeros((Q, dim), np.float32), 0.
		if normals is None or normal_ref is None:
			normals, normal_ref, weight_boundary = np.zeros((Q, dim), np.float32), np.zeros(Q, np.float32), 0.
		direct = direct if direct is not None else self.zero_grads()
		w = np.array([weight, weight_boundary], np.float64)
		sg = None if stop_gradient is None else np.ascontiguousarray(stop_gradient, dtype=np.int32)
		val = np.ascontiguousarray(val, dtype=self.real)
		self._fn('backward_val')(C.byref(self._G), _p(self.positions), _p(self.scalings), _p(self.rotations), _p(self.values),
								 C.c_double(self.tau), C.c_int(dim), _p(x), C.c_long(Q), _p(val),
								 _p(_f32(ref)), _p(_f32(normals)), _p(_f32(normal_ref)), _p(w), _p(sg),
								 *[_p(a) for a in direct], C.c_int(self.nthreads))
		return direct

	def backward2d_grad(self, x, grad, ref_grad=None, weight_grad=0., ref_vor=None, weight_vor=0., weight_div=0., stop_gradient=None,
						direct=None, vor=None, div=None):
		"""loop 2 of get_grad_losses_ti (2D/GSR.py:396-476) with the wrapper's defaults (:485-509)."""
		assert self.D == 2
		x = _f32(x)
		Q, dim = x.shape[0], self.dim
		if ref_grad is None:
			ref_grad, weight_grad = np.zeros((Q, dim, 2), np.float32), 0.
		if ref_vor is None:
			ref_vor, weight_vor = np.zeros(Q, np.float32), 0.
		direct = direct if direct is not None else self.zero_grads()
		vor = vor if vor is not None else direct
		div = div if div is not None else direct
		w = np.array([weight_grad, weight_vor, weight_div], np.float64)
		sg = None if stop_gradient is None else np.ascontiguousarray(stop_gradient, dtype=np.int32)
		grad = np.ascontiguousarray(grad, dtype=self.real)
		self._fn('backward_grad')(C.byref(self._G), _p(self.positions), _p(self.scalings), _p(self.rotations), _p(self.values),
								  C.c_double(self.tau), C.c_int(dim), _p(x), C.c_long(Q), _p(grad),
								  _p(_f32(ref_grad)), _p(_f32(ref_vor)), _p(w), _p(sg),
								  *[_p(a) for a in direct], *[_p(a) for a in vor], *[_p(a) for a in div], C.c_int(self.nthreads))
		return direct, vor, div

	def rk4(self, start, dt, pos_only=True):
		"""advection_rk4 (3D/GSR.py:634-677; 2D/GSR.py:549-592)."""
		start = _f32(start)
		Q, D = start.shape[0], self.D
		goal = np.zeros((Q, D), self.real)
		deform = None if pos_only else np.zeros((Q, D, D), self.real)
		gval = None if pos_only else np.zeros((Q, D), self.real)
		ggrad = None if pos_only else np.zeros((Q, D, D), self.real)
		self._fn('rk4')(C.byref(self._G), _p(self.positions), _p(self.scalings), _p(self.rotations), _p(self.values),
						C.c_double(self.tau), _p(start), C.c_long(Q), C.c_double(dt), _p(goal), _p(deform), _p(gval), _p(ggrad), C.c_int(self.nthreads))
		return goal if pos_only else (goal, deform, gval, ggrad)

	def rk4_eval_points(self, start, dt):
		"""the five points at which advection_rk4 evaluates the field (start, three stages, end point)"""
		x = np.asarray(start, np.float64)
		dt = float(np.float32(dt))
		pts, vs = [x], []
		for c in (.5, .5, 1.):
			v, _ = self.forward(pts[-1].astype(np.float32), need_grad=False)
			vs.append(v.astype(np.float64))
			pts.append(x + dt * c * vs[-1])
		v3, _ = self.forward(pts[-1].astype(np.float32), need_grad=False)
		pts.append(x + dt / 6. * (vs[0] + 2. * vs[1] + 2. * vs[2] + v3))
		return pts

	def mark_neighbors(self, x):
		x = _f32(x)
		mark = np.zeros(self.N, np.int32)
		fn = lib().o3_mark_neighbors if self.D == 3 else lib().o2_mark_neighbors
		fn.restype = None
		fn(C.byref(self._G), _p(self.positions), _p(x), C.c_long(x.shape[0]), _p(mark))
		return mark

	def count_candidates(self, x):
		"""C of SURVEY 8(d): candidate visits for one evaluation of the points x (3D)."""
		assert self.D == 3
		x = _f32(x)
		out = C.c_longlong(0)
		lib().o3_count_pairs.restype = None
		lib().o3_count_pairs(C.byref(self._G), _p(x), C.c_long(x.shape[0]), C.byref(out))
		return out.value

	def classify_pairs(self, x, rel_band=1e-4):
		"""(n_accepted, n_in_band) per sample; band = |q - q_max| <= rel_band*q_max (SURVEY 8c tolerance policy)."""
		x = _f32(x)
		Q = x.shape[0]
		n_acc, n_band = np.zeros(Q, np.int32), np.zeros(Q, np.int32)
		fn = lib().o3_classify_pairs if self.D == 3 else lib().o2_classify_pairs
		fn.restype = None
		fn(C.byref(self._G), _p(self.positions), _p(self.scalings), _p(self.rotations), C.c_double(self.tau), C.c_double(rel_band),
		   _p(x), C.c_long(Q), _p(n_acc), _p(n_band))
		return n_acc, n_band


def vortex_particles(x, x0, w, U, a, real=np.float64, need_val=True, need_grad=True, nthreads=0):
	"""regularised Biot-Savart sum of 3D/init_cond.py:122-145 over the particles (x0, w): returns (res (Q,3) | None, jac (Q,3,3) | None)"""
	x = _f32(x)
	x0, w = np.ascontiguousarray(x0, real), np.ascontiguousarray(w, real)
	Q, M = x.shape[0], x0.shape[0]
	res = np.zeros((Q, 3), real) if need_val else None
	jac = np.zeros((Q, 3, 3), real) if need_grad else None
	f = getattr(lib(), 'o3_vortex_particles' + ('_f32' if real == np.float32 else '_f64'))
	f.restype = None
	f(_p(x), C.c_long(Q), _p(x0), _p(w), C.c_long(M), C.c_float(np.float32(U)), C.c_float(np.float32(a)), _p(res), _p(jac), C.c_int(nthreads or os.cpu_count() or 1))
	return res, jac


def mesh_area_presum(vertices, faces):
	"""ti_get_tri_area (3D/mesh_sampler.py:12-21): inclusive prefix sums of the triangle areas, float32, serial order"""
	v, f = _f32(vertices), np.ascontiguousarray(faces, np.int32)
	out = np.zeros(f.shape[0], np.float32)
	lib().o3_mesh_area_presum.restype = None
	lib().o3_mesh_area_presum(_p(v), _p(f), C.c_long(f.shape[0]), _p(out))
	return out


def mesh_sample(uniforms, vertices, normals, faces, facenormals, presum):
	"""ti_sample (3D/mesh_sampler.py:60-88) for given uniforms (n,3): returns (data (n,3), normal (n,3)), float32"""
	u, v, nr = _f32(uniforms), _f32(vertices), _f32(normals)
	f, fn, ps = np.ascontiguousarray(faces, np.int32), np.ascontiguousarray(facenormals, np.int32), _f32(presum)
	n = u.shape[0]
	data, normal = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
	lib().o3_mesh_sample.restype = None
	lib().o3_mesh_sample(C.c_long(n), _p(u), _p(v), _p(nr), _p(f), _p(fn), _p(ps), C.c_long(f.shape[0]), _p(data), _p(normal))
	return data, normal


def interp_val(field, positions, domain, real=np.float64, nthreads=0):
	"""ti_get_interp_val (3D/advance_density.py:24-50): trilinear samples of field (nx,ny,nz) at positions (...,3) -> (...)"""
	f = np.ascontiguousarray(field, real)
	p = np.ascontiguousarray(positions, real)
	out = np.zeros(p.shape[:-1], real)
	dims = np.array(f.shape, np.int32)
	dom = np.array(domain, np.float32)
	fn = getattr(lib(), 'o3_interp_val' + ('_f32' if real == np.float32 else '_f64'))
	fn.restype = None
	fn(_p(f), _p(dims), _p(p), C.c_long(out.size), _p(dom), _p(out), C.c_int(nthreads or os.cpu_count() or 1))
	return out


# ---------------------------------------------------------------------------------------------
# per-timestep optimisation: project() of 3D/advance.py:183-287 and step() of 3D/GSR.py:144-152, :704-716
# ---------------------------------------------------------------------------------------------

def dense_torch_value_gradient(positions, scalings, rotations, values, x):
	"""
	The reference's DENSE representation (GaussianSplatting3D, 3D/GSR.py:93-130: every Gaussian at every point, no truncation, no
	hash) as torch ops on whatever device the tensors live on: Sigma^-1 = (R S)(R S)^T with S = diag(exp(s)) (:93-116),
	u = sum_i v_i exp(-1/2 d^T Sigma^-1 d) (:118-122) and grad u = -sum_i v_i g_i (Sigma^-1 d)^T (:124-130).  O(N Q) memory.
	Used as the CPU baseline "B2" of bench.py and as a tau = 0 cross-check.  Returns (u, grad u).
	"""
	import torch
	q = rotations / (rotations ** 2).sum(-1, keepdim=True) ** .5
	r, a, b, c = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
	R = torch.stack([1 - 2 * (b * b + c * c), 2 * (a * b - r * c), 2 * (a * c + r * b),
					 2 * (a * b + r * c), 1 - 2 * (a * a + c * c), 2 * (b * c - r * a),
					 2 * (a * c - r * b), 2 * (b * c + r * a), 1 - 2 * (a * a + b * b)], -1).reshape(-1, 3, 3)
	A = R @ torch.diag_embed(torch.exp(scalings))
	sigma_inv = A @ A.transpose(-1, -2)
	d = x[:, None, :] - positions[None, :, :]
	w = (sigma_inv[None] @ d[..., None]).squeeze(-1)
	per = values[None] * torch.exp(-.5 * (d * w).sum(-1))[..., None]
	grad = -(per[..., :, None] * w[..., None, :]).sum(1)
	return per.sum(1), grad


def split_gaussians(D, positions, scalings, rotations, values, normals, clamp_box=None):
	"""
	The reseeding split of clone_velocity_field, one round, in float64: 3D/advance.py:63-90 (axis ratio >= 2; children's scalings:
	the split axis += ln 2, every axis -= ln 2 / 3; children clamped to the extended domain) and 2D/advance.py:66-88 (ratio >= 1.5;
	the smaller of the two log inverse radii += ln 1.5; no clamp).  Child positions = mu + chol(Sigma) z with Sigma^-1 =
	R diag(e^{2s}) R^T symmetrised — what torch.distributions.MultivariateNormal(mu, precision_matrix=...).sample((2,)) returns for
	the standard-normal draws z = `normals` (2, n_split, D).  Returns (positions, scalings, rotations, values, stop_gradient) in the
	reference's layout: kept Gaussians first, then [first samples | second samples].
	"""
	P, S, R, V = [np.asarray(a, np.float64) for a in (positions, scalings, rotations, values)]
	ratio = np.exp(S.max(axis=1) - S.min(axis=1))
	split = ratio >= (2. if D == 3 else 1.5)
	ns = int(split.sum())
	keep = ~split
	if D == 3:
		q = R[split] / np.sqrt((R[split] ** 2).sum(axis=1, keepdims=True))
		r, a, b, c = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
		Rm = np.stack([1 - 2 * (b * b + c * c), 2 * (a * b - r * c), 2 * (a * c + r * b), 2 * (a * b + r * c), 1 - 2 * (a * a + c * c), 2 * (b * c - r * a),
					   2 * (a * c - r * b), 2 * (b * c + r * a), 1 - 2 * (a * a + b * b)], -1).reshape(-1, 3, 3)
	else:
		th = R[split].reshape(-1)
		Rm = np.stack([np.cos(th), -np.sin(th), np.sin(th), np.cos(th)], -1).reshape(-1, 2, 2)
	prec = np.einsum('nij,nj,nkj->nik', Rm, np.exp(2. * S[split]), Rm)
	prec = .5 * (prec + prec.transpose(0, 2, 1))
	L = np.linalg.cholesky(np.linalg.inv(prec)) if ns else np.zeros((0, D, D))
	z = np.asarray(normals, np.float64).reshape(2, ns, D)
	child = P[split][None] + np.einsum('nij,cnj->cni', L, z)
	Sc = S[split].copy()
	if D == 3:
		ax = Sc.argmin(axis=1)
		Sc[np.arange(ns), ax] += np.log(2.)
		Sc -= np.log(2.) / 3.
	else:
		ax = (Sc[:, 1] < Sc[:, 0]).astype(int)
		Sc[np.arange(ns), ax] += np.log(1.5)
	if clamp_box is not None:
		lo, hi = np.asarray(clamp_box[0::2], np.float64), np.asarray(clamp_box[1::2], np.float64)
		child = np.minimum(np.maximum(child, lo), hi)
	rep = lambda a: np.concatenate([a[keep], a[split], a[split]], axis=0)
	return (np.concatenate([P[keep], child.reshape(2 * ns, D)], axis=0), np.concatenate([S[keep], Sc, Sc], axis=0), rep(R), rep(V),
			np.concatenate([np.ones(int(keep.sum()), bool), np.zeros(2 * ns, bool)]))


class _Adam:
	"""torch.optim.Adam (defaults: betas .9/.999, eps 1e-8, no weight decay) + ReduceLROnPlateau(mode='min', threshold 1e-4 rel,
	cooldown 0, eps 1e-8) for one parameter tensor, in float64 (torch works in the tensor's float32)"""

	def __init__(self, lr, patience=50, factor=.9):
		self.lr, self.patience, self.factor = float(lr), patience, factor
		self.t, self.m, self.v = 0, None, None
		self.best, self.bad = np.inf, 0

	def step(self, p, g):
		if self.m is None:
			self.m, self.v = np.zeros_like(p, np.float64), np.zeros_like(p, np.float64)
		self.t += 1
		self.m += (g - self.m) * (1. - .9)
		self.v = self.v * .999 + (1. - .999) * g * g
		bc1, bc2 = 1. - .9 ** self.t, 1. - .999 ** self.t
		return p - (self.lr / bc1) * (self.m / (np.sqrt(self.v) / np.sqrt(bc2) + 1e-8))

	def schedule(self, metric):
		if metric < self.best * (1. - 1e-4):
			self.best, self.bad = metric, 0
		else:
			self.bad += 1
		if self.bad > self.patience:
			new_lr = max(self.lr * self.factor, 0.)
			if self.lr - new_lr > 1e-8:
				self.lr = new_lr
			self.bad = 0


def _curl3(J):
	return np.stack([J[:, 2, 1] - J[:, 1, 2], J[:, 0, 2] - J[:, 2, 0], J[:, 1, 0] - J[:, 0, 1]], axis=1)


class OracleProjector3D:
	"""
	One `project` phase of 3D/advance.py:183-287 on oracle fields: the advected-covector reference (:24-49) of the previous field,
	the vorticity / helicity / divergence losses with separate gradient sets and their PCGrad projection (:202-225), the anisotropy
	and volume regularisers (closed-form gradients; the reference uses autograd, :237-244), the boundary loss (:246-254), the
	scheduler metric (:256, without the helicity loss), 4 x Adam + 4 x ReduceLROnPlateau (3D/GSR.py:50-71, :144-152) and the grid
	rebuild with the new grid_scale (:704-716).  Parameters evolve in float64.
	"""
	LRS = (3e-4, 1e-5, 3e-4, 1e-5)	# positions, scalings, rotations, values (3D/advance.py:258-261)

	def __init__(self, bounds, params, previous, dt, boundary_lambda, tau, min_grid_scale, precision='f64', nthreads=1):
		self.bounds, self.dt, self.lam, self.tau, self.mgs = tuple(bounds), float(dt), float(boundary_lambda), float(tau), float(min_grid_scale)
		self.ext = extended_bounds(3, self.bounds, self.mgs)
		self.params = [np.array(p, np.float64) for p in params]
		self.previous = previous	# OracleGSR of the field that advects (static during the phase)
		self.prec, self.nthreads = precision, nthreads
		self.opt = [_Adam(lr, patience=50, factor=.9) for lr in self.LRS]
		self.grid_scale = None

	def _field(self):
		p = self.params
		return OracleGSR(3, self.ext, p[0], p[1], p[2], p[3], self.tau, self.mgs, precision=self.prec, nthreads=self.nthreads)

	def iterate(self, data, boundary=None):
		f = self._field()
		N = f.N
		x = np.asarray(data, np.float32)
		# reference: advected covector field (3D/advance.py:35-47)
		psi, dpsi, pb_v, pb_dv = self.previous.rk4(x, -self.dt, pos_only=False)
		pb_vor = _curl3(np.asarray(pb_dv, np.float64))
		ref_hel = (np.asarray(pb_v, np.float64) * pb_vor).sum(axis=1)
		ref_vor = np.einsum('qij,qj->qi', np.linalg.inv(np.asarray(dpsi, np.float64)), pb_vor)
		# losses with separate gradient sets (weights 1, 1, 1)
		val, grad = f.forward(x)
		direct, vor, div = f.zero_grads(), f.zero_grads(), f.zero_grads()
		f.backward3d(x, val, grad, ref_vor=ref_vor, weight_vor=1., ref_hel=ref_hel, weight_hel=1., weight_div=1., direct=direct, vor=vor, div=div)
		total = [np.asarray(d, np.float64) for d in direct]
		for k in range(4):	# PCGrad (:202-225)
			g1, g2 = np.asarray(vor[k], np.float64).copy(), np.asarray(div[k], np.float64).copy()
			if (g1 * g2).sum() < 0.:
				n1, n2 = g1 / np.sqrt((g1 ** 2).sum()), g2 / np.sqrt((g2 ** 2).sum())
				g1, g2 = g1 - (g1 * n2).sum() * n2, g2 - (g2 * n1).sum() * n1
			total[k] = total[k] + g1 + g2
		om = _curl3(np.asarray(grad, np.float64))
		loss_vor = np.abs(om - ref_vor).mean(axis=1).mean()
		loss_div = ((np.asarray(grad, np.float64)[:, 0, 0] + grad[:, 1, 1] + grad[:, 2, 2]) ** 2).mean()
		# regularisers (:237-244): 10 * aniso + 10 * vol (+ 0 * |values|)
		s = self.params[1]
		ratio = np.exp(s.max(axis=1) - s.min(axis=1))
		loss_aniso = (np.where(ratio >= 1.5, ratio, 1.5) - 1.5).mean()
		vol = np.exp(-s.sum(axis=1))
		r = vol / vol.mean()
		loss_vol = ((r - 1.) ** 2).mean()
		gs = np.zeros_like(s)
		kmax, kmin = s.argmax(axis=1), s.argmin(axis=1)	# first index on ties, like torch
		on = (ratio >= 1.5) & (kmax != kmin)
		idx = np.arange(N)
		gs[idx[on], kmax[on]] += 10. * ratio[on] / N
		gs[idx[on], kmin[on]] -= 10. * ratio[on] / N
		gs += (-10. * 2. / N * r * (r - (r ** 2).mean()))[:, None]	# d/ds_k of mean((V/mean V - 1)^2)
		total[1] = total[1] + gs
		# boundary loss (:246-254)
		boundary_constraint = 0.
		if self.lam and boundary is not None:
			bx, bn = np.asarray(boundary[0], np.float32), np.asarray(boundary[1], np.float32)
			bval, _ = f.forward(bx, need_grad=False)
			bd = f.zero_grads()
			f.backward3d(bx, bval, np.zeros((bx.shape[0], 3, 3), f.real), normals=bn, weight_boundary=self.lam, direct=bd)
			for k in range(4):
				total[k] = total[k] + np.asarray(bd[k], np.float64)
			boundary_constraint = np.abs((np.asarray(bval, np.float64) * bn).sum(axis=1)).mean()
		loss_tot = loss_vor + loss_div + 10. * loss_aniso + 10. * loss_vol + self.lam * boundary_constraint	# (:256: no helicity term)
		# what the step consumes (for the parity tests): the raw sets, the total gradient, the scheduler metric
		self.last = {'vor': [np.asarray(a, np.float64).copy() for a in vor], 'div': [np.asarray(a, np.float64).copy() for a in div],
					 'total': [np.asarray(a, np.float64).copy() for a in total], 'metric': float(loss_tot), 'lr': [o.lr for o in self.opt]}
		# step (3D/GSR.py:144-152, :714-716): Adam x4, schedulers x4 on the same metric, new grid
		for k in range(4):
			self.params[k] = self.opt[k].step(self.params[k], total[k])
			self.opt[k].schedule(float(np.float32(loss_tot)))
		self.grid_scale = grid_scale_of(self.tau, np.asarray(self.params[1], np.float32), self.mgs, self.ext)
		return loss_tot


class OracleProjector2D:
	"""
	One `project` phase of 2D/advance.py:186-291 on oracle fields: the advected vorticity of the previous field, zero outside the
	advance domain (:21-56), vorticity / divergence losses with separate gradient sets (get_grad_losses) and their PCGrad projection
	(:187-193, :226-233), value samples (boundary_generator_1 through get_losses, :221-224), normal samples (boundary_generator_2,
	:235-239), the anisotropy / volume / position-drift regularisers (:252-260; closed-form gradients), the scheduler metric (:262),
	4 x Adam + 4 x ReduceLROnPlateau at lr 1e-4 (:264-269) and the grid rebuild.  Parameters evolve in float64.
	"""

	def __init__(self, bounds, params, previous, dt, boundary_lambda, tau, min_grid_scale, domain, precision='f64', nthreads=1):
		self.bounds, self.dt, self.lam, self.tau, self.mgs = tuple(bounds), float(dt), float(boundary_lambda), float(tau), float(min_grid_scale)
		self.domain = tuple(domain)	# the advance domain in GSR space: back-traced points outside it carry no vorticity
		self.ext = extended_bounds(2, self.bounds, self.mgs)
		self.params = [np.array(p, np.float64) for p in params]
		self.positions_org = self.params[0].copy()
		self.previous = previous
		self.prec, self.nthreads = precision, nthreads
		self.opt = [_Adam(1e-4, patience=50, factor=.9) for _ in range(4)]
		self.grid_scale = None

	def _field(self):
		p = self.params
		return OracleGSR(2, self.ext, p[0], p[1], p[2], p[3], self.tau, self.mgs, precision=self.prec, nthreads=self.nthreads)

	def iterate(self, data, boundary_1=None, boundary_2=None):
		f = self._field()
		N = f.N
		x = np.asarray(data, np.float32)
		bk, _, _, dv = self.previous.rk4(x, -self.dt, pos_only=False)
		dv = np.asarray(dv, np.float64)
		ref_vor = dv[:, 1, 0] - dv[:, 0, 1]
		x0, x1, y0, y1 = self.domain
		bk = np.asarray(bk, np.float64)
		ref_vor[(bk[:, 0] < x0) | (bk[:, 0] > x1) | (bk[:, 1] < y0) | (bk[:, 1] > y1)] = 0.
		val, grad = f.forward(x)
		direct, vor, div = f.zero_grads(), f.zero_grads(), f.zero_grads()
		f.backward2d_grad(x, grad, ref_vor=ref_vor, weight_vor=1., weight_div=1., direct=direct, vor=vor, div=div)
		boundary_constraint = 0.
		if self.lam > 0. and boundary_1 is not None:
			bx, bv = np.asarray(boundary_1[0], np.float32), np.asarray(boundary_1[1], np.float32)
			out, _ = f.forward(bx, need_grad=False)
			f.backward2d_val(bx, out, ref=bv, weight=self.lam, direct=direct)
			boundary_constraint += np.abs(np.asarray(out, np.float64) - bv).mean()
		total = [np.asarray(d, np.float64).copy() for d in direct]
		for k in range(4):
			g1, g2 = np.asarray(vor[k], np.float64).copy(), np.asarray(div[k], np.float64).copy()
			if (g1 * g2).sum() < 0.:
				n1, n2 = g1 / np.sqrt((g1 ** 2).sum()), g2 / np.sqrt((g2 ** 2).sum())
				g1, g2 = g1 - (g1 * n2).sum() * n2, g2 - (g2 * n1).sum() * n1
			total[k] = total[k] + g1 + g2
		if self.lam > 0. and boundary_2 is not None:
			bx, bn, br = [np.asarray(a, np.float32) for a in boundary_2]
			out, _ = f.forward(bx, need_grad=False)
			bd = f.zero_grads()
			f.backward2d_val(bx, out, normals=bn, normal_ref=br, weight_boundary=self.lam, direct=bd)
			for k in range(4):
				total[k] = total[k] + np.asarray(bd[k], np.float64)
			boundary_constraint += np.abs((np.asarray(out, np.float64) * bn).sum(axis=1) - br).mean()
		grad = np.asarray(grad, np.float64)
		loss_vor = np.abs(grad[:, 1, 0] - grad[:, 0, 1] - ref_vor).mean()
		loss_div = ((grad[:, 0, 0] + grad[:, 1, 1]) ** 2).mean()
		s = self.params[1]
		ratio = np.exp(s.max(axis=1) - s.min(axis=1))
		loss_aniso = (np.where(ratio >= 1.5, ratio, 1.5) - 1.5).mean()
		vol = np.exp(-s.sum(axis=1))
		r = vol / vol.mean()
		loss_vol = ((r - 1.) ** 2).mean()
		gs = np.zeros_like(s)
		kmax, kmin = s.argmax(axis=1), s.argmin(axis=1)
		on = (ratio >= 1.5) & (kmax != kmin)
		idx = np.arange(N)
		gs[idx[on], kmax[on]] += 10. * ratio[on] / N
		gs[idx[on], kmin[on]] -= 10. * ratio[on] / N
		gs += (-10. * 2. / N * r * (r - (r ** 2).mean()))[:, None]
		total[1] = total[1] + gs
		dpos = self.params[0] - self.positions_org
		loss_delta_pos = (dpos ** 2).mean()
		total[0] = total[0] + .5 * 2. * dpos / dpos.size
		loss_tot = loss_vor + loss_div + 10. * loss_aniso + 10. * loss_vol + .5 * loss_delta_pos + self.lam * boundary_constraint
		self.last = {'vor': [np.asarray(a, np.float64).copy() for a in vor], 'div': [np.asarray(a, np.float64).copy() for a in div],
					 'total': [np.asarray(a, np.float64).copy() for a in total], 'metric': float(loss_tot), 'lr': [o.lr for o in self.opt]}
		for k in range(4):
			shape = self.params[k].shape
			self.params[k] = self.opt[k].step(self.params[k], total[k].reshape(shape))
			self.opt[k].schedule(float(np.float32(loss_tot)))
		self.grid_scale = grid_scale_of(self.tau, np.asarray(self.params[1], np.float32), self.mgs, self.ext)
		return loss_tot


class OracleFit3D:
	"""fit_velocity_with_gradient of 3D/initialize.py:9-46 on oracle fields: value + gradient L1 losses (one gradient set), the
	anisotropy and volume regularisers with weight 1, Adam x4 + ReduceLROnPlateau x4 at the class's learning rates, grid rebuild"""

	def __init__(self, bounds, params, lrs, tau, min_grid_scale, precision='f64', nthreads=1):
		self.bounds, self.tau, self.mgs = tuple(bounds), float(tau), float(min_grid_scale)
		self.ext = extended_bounds(3, self.bounds, self.mgs)
		self.params = [np.array(p, np.float64) for p in params]
		self.prec, self.nthreads = precision, nthreads
		self.opt = [_Adam(lr, patience=50, factor=.9) for lr in lrs]	# initialize_optimizers() (3D/GSR.py:50-71): factor .9, patience 50
		self.grid_scale = None

	def iterate(self, data, ref_val, ref_grad):
		p = self.params
		f = OracleGSR(3, self.ext, p[0], p[1], p[2], p[3], self.tau, self.mgs, precision=self.prec, nthreads=self.nthreads)
		N = f.N
		x = np.asarray(data, np.float32)
		val, grad = f.forward(x)
		direct = f.zero_grads()
		f.backward3d(x, val, grad, ref_val=ref_val, weight_val=1., ref_grad=ref_grad, weight_grad=1., direct=direct)
		total = [np.asarray(d, np.float64).copy() for d in direct]
		s = p[1]
		ratio = np.exp(s.max(axis=1) - s.min(axis=1))
		loss_aniso = (np.where(ratio >= 1.5, ratio, 1.5) - 1.5).mean()
		vol = np.exp(-s.sum(axis=1))
		r = vol / vol.mean()
		loss_vol = ((r - 1.) ** 2).mean()
		gs = np.zeros_like(s)
		kmax, kmin = s.argmax(axis=1), s.argmin(axis=1)
		on = (ratio >= 1.5) & (kmax != kmin)
		idx = np.arange(N)
		gs[idx[on], kmax[on]] += ratio[on] / N
		gs[idx[on], kmin[on]] -= ratio[on] / N
		gs += (-2. / N * r * (r - (r ** 2).mean()))[:, None]
		total[1] = total[1] + gs
		loss = np.abs(np.asarray(val, np.float64) - ref_val).mean() + np.abs(np.asarray(grad, np.float64) - ref_grad).mean() + loss_aniso + loss_vol
		self.last = {'total': [np.asarray(a, np.float64).copy() for a in total], 'metric': float(loss), 'lr': [o.lr for o in self.opt]}
		for k in range(4):
			self.params[k] = self.opt[k].step(self.params[k], total[k])
			self.opt[k].schedule(float(np.float32(loss)))
		self.grid_scale = grid_scale_of(self.tau, np.asarray(self.params[1], np.float32), self.mgs, self.ext)
		return loss


class OracleFit2D:
	"""fit_velocity_with_gradient of 2D/initialize.py:10-41 on oracle fields: value loss (get_losses, weight 1) and gradient loss
	(get_grad_losses, weight_grad 1) into one gradient set, anisotropy and volume regularisers with weight 1, Adam x4 +
	ReduceLROnPlateau x4 (factor .9, patience 50) at the learning rates set by the caller, grid rebuild"""

	def __init__(self, bounds, params, lrs, tau, min_grid_scale, precision='f64', nthreads=1):
		self.bounds, self.tau, self.mgs = tuple(bounds), float(tau), float(min_grid_scale)
		self.ext = extended_bounds(2, self.bounds, self.mgs)
		self.params = [np.array(p, np.float64) for p in params]
		self.prec, self.nthreads = precision, nthreads
		self.opt = [_Adam(lr, patience=50, factor=.9) for lr in lrs]
		self.grid_scale = None

	def iterate(self, data, ref_val, ref_grad):
		p = self.params
		f = OracleGSR(2, self.ext, p[0], p[1], p[2], p[3], self.tau, self.mgs, precision=self.prec, nthreads=self.nthreads)
		N = f.N
		x = np.asarray(data, np.float32)
		val, grad = f.forward(x)
		direct = f.zero_grads()
		f.backward2d_val(x, val, ref=ref_val, weight=1., direct=direct)
		f.backward2d_grad(x, grad, ref_grad=ref_grad, weight_grad=1., direct=direct)
		total = [np.asarray(d, np.float64).copy() for d in direct]
		s = p[1]
		ratio = np.exp(s.max(axis=1) - s.min(axis=1))
		loss_aniso = (np.where(ratio >= 1.5, ratio, 1.5) - 1.5).mean()
		vol = np.exp(-s.sum(axis=1))
		r = vol / vol.mean()
		loss_vol = ((r - 1.) ** 2).mean()
		gs = np.zeros_like(s)
		kmax, kmin = s.argmax(axis=1), s.argmin(axis=1)
		on = (ratio >= 1.5) & (kmax != kmin)
		idx = np.arange(N)
		gs[idx[on], kmax[on]] += ratio[on] / N
		gs[idx[on], kmin[on]] -= ratio[on] / N
		gs += (-2. / N * r * (r - (r ** 2).mean()))[:, None]
		total[1] = total[1] + gs
		loss = np.abs(np.asarray(val, np.float64) - ref_val).mean() + np.abs(np.asarray(grad, np.float64) - ref_grad).mean() + loss_aniso + loss_vol
		self.last = {'total': [np.asarray(a, np.float64).copy() for a in total], 'metric': float(loss), 'lr': [o.lr for o in self.opt]}
		for k in range(4):
			shape = self.params[k].shape
			self.params[k] = self.opt[k].step(self.params[k], total[k].reshape(shape))
			self.opt[k].schedule(float(np.float32(loss)))
		self.grid_scale = grid_scale_of(self.tau, np.asarray(self.params[1], np.float32), self.mgs, self.ext)
		return loss
