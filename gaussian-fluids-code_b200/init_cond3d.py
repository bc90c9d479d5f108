"""
Scenario tables and boundary samplers of the reference's 3D/init_cond.py (domains :13-32, ring parameters :39-108,
box sampler :227-249).  The analytic Biot-Savart fields (:115-216) are evaluated with plain torch here (they run once
per initial fit, not per time step — SURVEY 8f row N2); the OBJ mesh sampler (:223-226) is SURVEY row N3.
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


def _ring_particles(ring):
	"""n vortex particles on the ring and their (strength-scaled) tangents (3D/init_cond.py:147-156)"""
	device = gsr3d.device
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
	return x0, w, ring['radius'] / (2 * ring['n']), ring['thickness']


def _chunks(x, size=16384):
	for b in range(0, x.shape[0], size):
		yield slice(b, min(b + size, x.shape[0]))


def vortex_ring(x, ring):
	"""regularised Biot-Savart velocity of one ring: sum_j U f(r) (w_j x d),  f = (1 - exp(-(r/a)^3)) / r^3 (3D/init_cond.py:122-131)"""
	x0, w, U, a = _ring_particles(ring)
	res = torch.zeros_like(x)
	for sl in _chunks(x):
		d = x[sl, None, :] - x0[None, :, :]
		r = d.norm(dim=-1)
		fr = (1. - torch.exp(-(r / a) ** 3)) / r ** 3
		res[sl] = (U * fr[..., None] * torch.linalg.cross(w[None].expand_as(d), d)).sum(dim=1)
	return res


def vortex_ring_gradient(x, ring):
	"""Jacobian of vortex_ring (3D/init_cond.py:132-145): U (f'/r) [w]x d d^T + U f [w]x"""
	x0, w, U, a = _ring_particles(ring)
	W = torch.zeros((w.shape[0], 3, 3), device=x.device)
	W[:, 0, 1], W[:, 0, 2], W[:, 1, 0], W[:, 1, 2], W[:, 2, 0], W[:, 2, 1] = -w[:, 2], w[:, 1], w[:, 2], -w[:, 0], -w[:, 1], w[:, 0]
	res = torch.zeros((x.shape[0], 3, 3), device=x.device)
	for sl in _chunks(x, 4096):
		d = x[sl, None, :] - x0[None, :, :]
		r = d.norm(dim=-1)
		ex = torch.exp(-(r / a) ** 3)
		fr = (1. - ex) / r ** 3
		frp = -3. / r ** 4 * (1. - ex) + 3. / (a ** 3 * r) * ex
		Wd = torch.einsum('jkl,qjl->qjk', W, d)
		res[sl] = (U * (frp / r)[..., None, None] * Wd[..., :, None] * d[..., None, :] + U * fr[..., None, None] * W[None]).sum(dim=1)
	return res


def make_field(init_cond):
	"""velocity field callable with a .gradient attribute, like `eval(cmd_args.init_cond)` in 3D/initialize.py:53"""
	rings = rings_of(init_cond)

	def field(x):
		return sum(vortex_ring(x, r) for r in rings)
	field.gradient = lambda x: sum(vortex_ring_gradient(x, r) for r in rings)
	return field


def sample_on_box(n, x_min, x_max, y_min, y_max, z_min, z_max):
	"""
	n points on the faces of the box, area-weighted, with inward normals (3D/init_cond.py:227-249).
	Written without boolean-mask indexing so that it never synchronises the host (the reference's version does).
	"""
	device = gsr3d.device
	sx, sy, sz = x_max - x_min, y_max - y_min, z_max - z_min
	areas = torch.tensor([sy * sz, sy * sz, sz * sx, sz * sx, sx * sy, sx * sy], device=device)
	t = torch.rand(n, device=device) * areas.sum()
	face = torch.bucketize(t, torch.cumsum(areas, 0)[:-1], right=True)	# 0..5: x_min, x_max, y_min, y_max, z_min, z_max
	uvw = torch.rand((n, 3), device=device) * torch.tensor([sx, sy, sz], device=device) + torch.tensor([x_min, y_min, z_min], device=device)
	axis = face // 2
	upper = (face % 2).to(torch.float32)
	lo = torch.tensor([x_min, y_min, z_min], device=device)[axis]
	hi = torch.tensor([x_max, y_max, z_max], device=device)[axis]
	onehot = torch.nn.functional.one_hot(axis, 3).to(torch.float32)
	data = uvw * (1. - onehot) + onehot * (lo + upper * (hi - lo))[:, None]
	normal = onehot * (1. - 2. * upper)[:, None]
	return data.contiguous(), normal.contiguous()


def make_boundary_sampler(init_cond):
	x_min, x_max, y_min, y_max, z_min, z_max = domain[init_cond]
	if init_cond == 'ring_with_obstacle':
		raise NotImplementedError('mesh boundary sampler (assets/bunny.obj is not shipped with the reference): SURVEY 8f row N3')
	return lambda n: sample_on_box(n, x_min, x_max, y_min, y_max, z_min, z_max)
