"""
Scenario tables and boundary samplers of the reference's 3D/init_cond.py (domains :13-32, ring parameters :39-108,
box sampler :227-249) and the analytic Biot-Savart fields (:115-216) on the CUDA kernel of csrc/fields.cu;
the OBJ mesh sampler (:223-226) is SURVEY row N3.
"""
import torch

from . import gsr3d

# [x_min, x_max, y_min, y_max, z_min, z_max]
domain = {name: (0., 1., 0., 1., 0., 1.) for name in ('leapfrog', 'single_vortex_ring', 'ring_collide', 'ring_with_obstacle')}
initial_particle_count = {'leapfrog': (10, 10, 10), 'single_vortex_ring': (40, 40, 40), 'ring_collide': (40, 40, 40), 'ring_with_obstacle': (40, 40, 40)}
visualize_res = {name: (128, 128, 128) for name in domain}


def _ring(center, normal, radius, thickness, strength, n=500):
	return dict(center=center, normal=normal, radius=radius, thickness=thickness, strength=strength, n=n)


other_info = {
	'leapfrog': {'ring1': _ring([.75, .5, .5], [-1., 0., 0.], 1. / 6, .12 / 6, .1 / 6), 'ring2': _ring([.85, .5, .5], [-1., 0., 0.], .7 / 6, .12 / 6, .1 / 6)},
	'single_vortex_ring': _ring([.5, .5, .5], [1., 0., 0.], 1. / 6, .1 / 6, .1 / 6),
	'ring_collide': {'ring1': _ring([-.5 / 6 + .5, .5, .5], [1., 0., 0.], .3 / 6, .12 / 6, .1 / 6), 'ring2': _ring([.5 / 6 + .5, .5, .5], [-1., 0., 0.], .3 / 6, .12 / 6, .1 / 6)},
	'ring_with_obstacle': {'obj_file': '../assets/bunny.obj', 'scale': 1. / 4.8, 'translate': [0.8225, 0.3150, 0.2650],
						   'rings': [_ring([.475, .6, .53], [.2 / 1.08, .2 / 1.08, -1. / 1.08], .05, .02, .2 / 6),
									 _ring([0.4380, 0.5630, 0.7152], [.2 / 1.08, .2 / 1.08, -1. / 1.08], .05, .02, .2 / 6)]},
}


def rings_of(init_cond):
	info = other_info[init_cond]
	if 'rings' in info:
		return info['rings']
	if 'ring1' in info:
		return [info['ring1'], info['ring2']]
	return [info]


_RING_CACHE = {}


def _ring_particles(ring):
	"""n vortex particles on the ring and their (strength-scaled) tangents (3D/init_cond.py:147-156); built once per ring and device
	(the construction uploads constants and reads a norm back: neither belongs into every evaluation, nor into a captured graph)"""
	device = gsr3d.device
	key = (id(ring), str(device))
	hit = _RING_CACHE.get(key)
	if hit is not None and hit[0] is ring:
		return hit[1]
	out = _ring_particles_build(ring, device)
	_RING_CACHE[key] = (ring, out)
	return out


def _ring_particles_build(ring, device):
	normal = torch.tensor(ring['normal'], device=device)
	center = torch.tensor(ring['center'], device=device)
	axis_x = torch.tensor([1., 0., 0.], device=device)
	if torch.linalg.cross(axis_x, normal).norm() < 1e-5:
		axis_x = torch.tensor([0., 1., 0.], device=device)
	axis_y = torch.linalg.cross(normal, axis_x)
	axis_y = axis_y / axis_y.norm()
	axis_x = torch.linalg.cross(axis_y, normal)
	theta = torch.linspace(0., 2. * torch.pi, ring['n'] + 1, device=device)[:-1]
	x0 = (axis_x[None] * torch.cos(theta)[:, None] + axis_y[None] * torch.sin(theta)[:, None]) * ring['radius'] + center
	w = (axis_x[None] * -torch.sin(theta)[:, None] + axis_y[None] * torch.cos(theta)[:, None]) * ring['strength']
	return x0.contiguous(), w.contiguous(), ring['radius'] / (2 * ring['n']), ring['thickness']


def _biot_savart(x, ring, val, grad):
	"""gsr_vortex_particles: val (Q,3) += velocity, grad (Q,3,3) += Jacobian of one ring's regularised Biot-Savart field"""
	import ctypes as C
	from . import _lib
	from ._lib import check, ptr, stream
	if not x.is_cuda:
		raise _lib.GsrError('the analytic initial fields run on the GPU (gsr_vortex_particles); got a CPU tensor')
	x0, w, U, a = _ring_particles(ring)
	x = x.detach().contiguous().float()
	check(_lib.lib().gsr_vortex_particles(ptr(x, name='x'), C.c_int64(x.shape[0]), ptr(x0.contiguous()), ptr(w.contiguous()), C.c_int64(x0.shape[0]),
										  C.c_float(U), C.c_float(a), ptr(val, allow_none=True), ptr(grad, allow_none=True), stream()), 'gsr_vortex_particles')


def vortex_ring(x, ring):
	"""regularised Biot-Savart velocity of one ring: sum_j U f(r) (w_j x d),  f = (1 - exp(-(r/a)^3)) / r^3 (3D/init_cond.py:122-131, :147-158)"""
	res = torch.zeros((x.shape[0], 3), dtype=torch.float32, device=x.device)
	_biot_savart(x, ring, res, None)
	return res


def vortex_ring_gradient(x, ring):
	"""Jacobian of vortex_ring (3D/init_cond.py:132-145, :159-170): U (f'/r) [w]x d d^T + U f [w]x"""
	res = torch.zeros((x.shape[0], 3, 3), dtype=torch.float32, device=x.device)
	_biot_savart(x, ring, None, res)
	return res


def make_field(init_cond):
	"""velocity field callable with a .gradient attribute, like `eval(cmd_args.init_cond)` in 3D/initialize.py:53; `.both(x)` returns
	(velocity, Jacobian) from one pass over the particles (the fit needs both for every batch)"""
	rings = rings_of(init_cond)

	def run(x, need_val, need_grad):
		val = torch.zeros((x.shape[0], 3), dtype=torch.float32, device=x.device) if need_val else None
		grad = torch.zeros((x.shape[0], 3, 3), dtype=torch.float32, device=x.device) if need_grad else None
		for r in rings:	# the kernel accumulates, like the reference's `res +=`
			_biot_savart(x, r, val, grad)
		return val, grad

	def field(x):
		return run(x, True, False)[0]
	field.gradient = lambda x: run(x, False, True)[1]
	field.both = lambda x: run(x, True, True)
	field.graph_safe = True	# a pure device function of its argument: the fit may replay it from a CUDA graph (graphloop.py)
	return field


def sample_on_box(n, x_min, x_max, y_min, y_max, z_min, z_max):
	"""
	n points on the faces of the box, area-weighted, with inward normals (3D/init_cond.py:227-249).
	Written without boolean-mask indexing so that it never synchronises the host (the reference's version does).
	"""
	from .init_cond2d import _const	# constants uploaded once: no host->device copy per call, capturable into a CUDA graph
	device = gsr3d.device
	sx, sy, sz = x_max - x_min, y_max - y_min, z_max - z_min
	areas = _const([sy * sz, sy * sz, sz * sx, sz * sx, sx * sy, sx * sy], device)
	t = torch.rand(n, device=device) * areas.sum()
	face = torch.bucketize(t, torch.cumsum(areas, 0)[:-1], right=True)	# 0..5: x_min, x_max, y_min, y_max, z_min, z_max
	uvw = torch.rand((n, 3), device=device) * _const([sx, sy, sz], device) + _const([x_min, y_min, z_min], device)
	axis = face // 2
	upper = (face % 2).to(torch.float32)
	lo = _const([x_min, y_min, z_min], device)[axis]
	hi = _const([x_max, y_max, z_max], device)[axis]
	onehot = torch.nn.functional.one_hot(axis, 3).to(torch.float32)
	data = uvw * (1. - onehot) + onehot * (lo + upper * (hi - lo))[:, None]
	normal = onehot * (1. - 2. * upper)[:, None]
	return data.contiguous(), normal.contiguous()


def make_boundary_sampler(init_cond, obj_file=None):
	"""`boundary_sampler[init_cond]` of 3D/init_cond.py:251-265; ring_with_obstacle adds n points on the obstacle mesh to the n
	points on the domain box (sample_for_ring_with_obstacle).  obj_file overrides the scene's asset path (the reference's
	assets/bunny.obj is not shipped with it)."""
	x_min, x_max, y_min, y_max, z_min, z_max = domain[init_cond]
	box = lambda n: sample_on_box(n, x_min, x_max, y_min, y_max, z_min, z_max)
	box.graph_safe = True	# pure device functions of torch's CUDA random stream / a device-resident draw counter (graphloop.py)
	info = other_info[init_cond]
	if 'obj_file' not in info:
		return box
	from .mesh_sampler import MeshSampler
	mesh = MeshSampler(obj_file or info['obj_file'], info['scale'], info.get('rotate', torch.eye(3)), info['translate'])

	def both(n):
		d1, n1 = box(n)
		d2, n2 = mesh.sample(n)
		return torch.cat([d1, d2], dim=0), torch.cat([n1, n2], dim=0)
	both.mesh = mesh
	both.graph_safe = True
	return both
