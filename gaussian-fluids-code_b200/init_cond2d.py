"""
2D scenarios — the data tables, analytic fields, boundary samplers and scale converters of the reference's
2D/init_cond.py, as an object (`Scene2D(init_cond)`) instead of module globals keyed by the command line.

Covered: taylor_green, taylor_vortex, leapfrog (closed-form fields, box boundary).  The obstacle scenes
(vortices_pass*, karman) need the circle / moving-inlet samplers of 2D/init_cond.py:267-428 and are not built yet.
"""
import numpy as np
import torch

from . import gsr2d

# 2D/init_cond.py:12-69
initialize_domain = {'taylor_green': (0., 2. * np.pi, 0., 2. * np.pi), 'taylor_vortex': (-5., 5., -5., 5.), 'leapfrog': (-5., 5., -5., 5.)}
advance_domain = dict(initialize_domain)
visualize_domain = dict(initialize_domain)
initial_particle_count = {'taylor_green': (24, 24), 'taylor_vortex': (71, 71), 'leapfrog': (71, 71)}
visualize_res = {'taylor_green': (200, 200), 'taylor_vortex': (200, 200), 'leapfrog': (200, 200)}
# 2D/init_cond.py:76-131
other_info = {
	'taylor_green': {},
	'taylor_vortex': {'U': 3., 'a': .5, 'vortex_pos1': (-.8, 0.), 'vortex_pos2': (.8, 0.)},
	'leapfrog': {'U': .5, 'a': .3, 'vortex_pos1': (-3., -3.), 'vortex_pos2': (-1., -3.), 'vortex_pos3': (1., -3.), 'vortex_pos4': (3., -3.)},
}


def _dev():
	return gsr2d.device


def vortex_particle(x, x0, radius, magnitude, grad):
	"""regularised point vortex and its Jacobian (2D/init_cond.py:138-156)"""
	eps = 1e-6
	dx = x - x0
	r = (dx ** 2).sum(dim=-1) ** .5
	ex = torch.exp(-((r + eps) / radius) ** 2)
	if not grad:
		return (magnitude * (r + eps) ** -2. * (1. - ex))[:, None] * torch.stack([-dx[:, 1], dx[:, 0]], dim=-1)
	p1 = torch.stack([dx[:, 0] * dx[:, 1], dx[:, 1] ** 2, -dx[:, 0] ** 2, -dx[:, 0] * dx[:, 1]], dim=-1).reshape(-1, 2, 2)
	p1 = p1 * (2. * magnitude / r / (r + eps) * ((r + eps) ** -2. * (1. - ex) - radius ** -2. * ex))[:, None, None]
	p2 = torch.zeros((x.shape[0], 2, 2), device=x.device)
	p2[:, 0, 1], p2[:, 1, 0] = -1., 1.
	return p1 + p2 * (magnitude * (r + eps) ** -2. * (1. - ex))[:, None, None]


def taylor_green(x, grad):
	"""2D/init_cond.py:158-167 — a steady solution of the Euler equations"""
	s0, c0, s1, c1 = torch.sin(x[:, 0]), torch.cos(x[:, 0]), torch.sin(x[:, 1]), torch.cos(x[:, 1])
	if grad:
		return torch.stack([c0 * c1, -s0 * s1, s0 * s1, -c0 * c1], dim=-1).reshape(-1, 2, 2)
	return torch.stack([s0 * c1, -c0 * s1], dim=1)


def taylor_vortex(x, grad):
	"""2D/init_cond.py:169-191"""
	info = other_info['taylor_vortex']
	U, a = info['U'], info['a']
	res = 0.
	for (cx, cy) in (info['vortex_pos1'], info['vortex_pos2']):
		r2 = (x[:, 0] - cx) ** 2 + (x[:, 1] - cy) ** 2
		amp = U / a * torch.exp(.5 * (1. - r2 / a ** 2))
		if grad:
			g = torch.stack([(cx - x[:, 0]) * (cy - x[:, 1]) / a ** 2, (cy - x[:, 1]) ** 2 / a ** 2 - 1.,
							 1. - (cx - x[:, 0]) ** 2 / a ** 2, (x[:, 0] - cx) * (cy - x[:, 1]) / a ** 2], dim=-1).reshape(-1, 2, 2)
			res = res + g * amp[:, None, None]
		else:
			res = res + torch.stack([cy - x[:, 1], x[:, 0] - cx], dim=1) * amp[:, None]
	return res


def leapfrog(x, grad):
	"""2D/init_cond.py:193-202: two co-rotating pairs"""
	info = other_info['leapfrog']
	U, a = info['U'], info['a']
	res = 0.
	for k, sign in ((1, 1.), (2, 1.), (3, -1.), (4, -1.)):
		res = res + vortex_particle(x, torch.tensor(info[f'vortex_pos{k}'], device=x.device), a, sign * U, grad)
	return res


FIELDS = {'taylor_green': taylor_green, 'taylor_vortex': taylor_vortex, 'leapfrog': leapfrog}


class Scene2D:
	"""everything 2D/init_cond.py derives from `--init_cond`"""

	def __init__(self, init_cond):
		if init_cond not in FIELDS:
			raise NotImplementedError(f'2D scene {init_cond!r} (obstacle scenes need the circle / inlet samplers of 2D/init_cond.py:267-428)')
		self.name = init_cond
		self.initialize_domain = initialize_domain[init_cond]
		self.advance_domain = advance_domain[init_cond]
		self.visualize_domain = visualize_domain[init_cond]
		self.particle_count = initial_particle_count[init_cond]
		self.visualize_res = visualize_res[init_cond]
		x_min, x_max, y_min, y_max = self.initialize_domain
		self.scaling_factor = 10. / min(x_max - x_min, y_max - y_min)	# 2D/init_cond.py:22-25
		self._field = FIELDS[init_cond]

	# ---- fields in "original" coordinates and their GSR-space ("target") versions (2D/init_cond.py:435-453) ----
	def velocity(self, x):
		return self._field(x, False)

	def gradient(self, x):
		return self._field(x, True)

	def target_velocity(self, x):
		return self.scaling_factor * self._field(x / self.scaling_factor, False)

	def target_gradient(self, x):
		return self._field(x / self.scaling_factor, True)

	def scaled(self, dom):
		return tuple(v * self.scaling_factor for v in dom)

	# ---- samplers -------------------------------------------------------------------------------------------------
	def data_generator(self, gaussian_splatting):
		"""default_data_generator of 2D/advance.py:314-316 / initialize.py: Q = N uniform samples of the advance domain, GSR space"""
		x_min, x_max, y_min, y_max = self.advance_domain
		dev = _dev()
		return (torch.rand_like(gaussian_splatting.positions.detach(), device=dev) * torch.tensor([x_max - x_min, y_max - y_min], device=dev)
				+ torch.tensor([x_min, y_min], device=dev)) * self.scaling_factor

	def test_generator(self):
		x_min, x_max, y_min, y_max = self.advance_domain
		return gsr2d.get_grid_points(x_min, x_max, y_min, y_max, *self.visualize_res) * self.scaling_factor

	def boundary_sampler_2(self, n):
		"""sample_on_domain_boundary_2 (2D/init_cond.py:306-325) in GSR space: points on the four edges (perimeter-weighted), OUTWARD
		normals, target normal velocity 0.  Written without boolean-mask indexing (no host sync)."""
		x_min, x_max, y_min, y_max = self.advance_domain
		xs, ys = x_max - x_min, y_max - y_min
		dev = _dev()
		t = torch.rand(n, device=dev) * (xs + ys) * 2.
		edge = (t >= xs).long() + (t >= xs + ys).long() + (t >= 2. * xs + ys).long()
		px = torch.stack([x_min + t, torch.full_like(t, x_max), x_max - t + xs + ys, torch.full_like(t, x_min)], dim=1)
		py = torch.stack([torch.full_like(t, y_min), y_min + t - xs, torch.full_like(t, y_max), y_max - t + 2. * xs + ys], dim=1)
		data = torch.stack([px.gather(1, edge[:, None])[:, 0], py.gather(1, edge[:, None])[:, 0]], dim=1)
		normals = torch.tensor([[0., -1.], [1., 0.], [0., 1.], [-1., 0.]], device=dev)[edge]
		return (data * self.scaling_factor).contiguous(), normals.contiguous(), torch.zeros(n, device=dev)

	@property
	def boundary_samplers(self):
		"""[boundary_generator_1, boundary_generator_2] as in 2D/init_cond.py:419-428"""
		return [None, self.boundary_sampler_2]
