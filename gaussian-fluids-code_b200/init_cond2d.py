"""
2D scenarios — the data tables, analytic fields, boundary samplers, scale converters and the moving inlet of the reference's
2D/init_cond.py, as an object (`Scene2D(init_cond)`) instead of module globals keyed by the command line.

All eight scenes of the reference: taylor_green, taylor_vortex, leapfrog (closed-form fields, box boundary); vortices_pass,
vortices_pass_narrow (two discs, free-slip), vortices_pass_noslip (two discs, no-slip: value samples on the discs), karman
(disc in a channel whose inlet moves with the flow: `extra_advector` / `extra_loader`), vortices_pass_particles (field from
an OBJ list of vortex particles — the asset is not shipped with the reference; pass `particles_obj=`).
Pinned by tests/golden/ref2d_scenes.npz, produced by the reference module itself (tests/golden/make_golden_scenes2d.py).
"""
import numpy as np
import torch

from . import gsr2d

# 2D/init_cond.py:12-69
initialize_domain = {'taylor_green': (0., 2. * np.pi, 0., 2. * np.pi), 'taylor_vortex': (-5., 5., -5., 5.), 'leapfrog': (-5., 5., -5., 5.),
					 'vortices_pass': (0., 1., 0., 1.), 'vortices_pass_narrow': (0., 1., 0., 1.), 'vortices_pass_noslip': (0., 1., 0., 1.),
					 'vortices_pass_particles': (-5., 5., -5., 5.), 'karman': (-6.10321, 1.906778, -0.598466, 0.60349)}
advance_domain = dict(initialize_domain)
visualize_domain = dict(initialize_domain, vortices_pass_particles=(-2.5, 2.5, -2.5, 2.5), karman=(-1.10321, 1.906778, -0.598466, 0.60349))
initial_particle_count = dict({k: (71, 71) for k in initialize_domain}, taylor_green=(24, 24), karman=(400, 60))
visualize_res = dict({k: (200, 200) for k in initialize_domain}, karman=(501, 200))
# 2D/init_cond.py:76-131
_PASS = {'U': 5e-3, 'a': 3e-2, 'vortex_pos1': (.1, .525), 'vortex_pos2': (.1, .475), 'obstacle_radius': 60. / 511.}
other_info = {
	'taylor_green': {},
	'taylor_vortex': {'U': 3., 'a': .5, 'vortex_pos1': (-.8, 0.), 'vortex_pos2': (.8, 0.)},
	'leapfrog': {'U': .5, 'a': .3, 'vortex_pos1': (-3., -3.), 'vortex_pos2': (-1., -3.), 'vortex_pos3': (1., -3.), 'vortex_pos4': (3., -3.)},
	'vortices_pass': dict(_PASS, obstacle_pos1=(.5, .27), obstacle_pos2=(.5, .73)),
	'vortices_pass_narrow': dict(_PASS, obstacle_pos1=(.5, .285), obstacle_pos2=(.5, .715)),
	'vortices_pass_noslip': dict(_PASS, obstacle_pos1=(.5, .27), obstacle_pos2=(.5, .73)),
	'vortices_pass_particles': {'particles_obj': '../assets/vortices_pass_particles.obj', 'obstacle_pos1': (0., 1.), 'obstacle_pos2': (0., -1.), 'obstacle_radius': .25},
	'karman': {'v_magnitude': .5, 'obstacle_pos': (-0.80356845, -0.00502235), 'obstacle_radius': 0.04553178393357534, 'd0': np.pi / 15.},
}


def _dev():
	return gsr2d.device


_CONSTS = {}


def _const(values, device):
	"""small constant tensors, uploaded once: a host->device copy per call is a synchronising pageable copy, and is not allowed
	while the optimisation loops are being captured into a CUDA graph (graphloop.py)"""
	import numpy as _np
	a = _np.asarray(values, dtype=_np.float32)
	key = (a.tobytes(), a.shape, str(device))
	t = _CONSTS.get(key)
	if t is None:
		if len(_CONSTS) > 4096:
			_CONSTS.clear()
		t = _CONSTS[key] = torch.tensor(a, device=device)
	return t


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
		res = res + vortex_particle(x, _const(info[f'vortex_pos{k}'], x.device), a, sign * U, grad)
	return res


def vortices_pass_of(name):
	"""2D/init_cond.py:204-211: a counter-rotating pair of regularised point vortices"""
	info = other_info[name]

	def field(x, grad):
		return vortex_particle(x, _const(info['vortex_pos1'], x.device), info['a'], info['U'], grad) \
			+ vortex_particle(x, _const(info['vortex_pos2'], x.device), info['a'], -info['U'], grad)
	return field


def karman(x, grad):
	"""2D/init_cond.py:251-260: uniform inflow (v_magnitude, 0); its Jacobian is zero"""
	if grad:
		return torch.zeros((x.shape[0], 2, 2), device=x.device)
	res = torch.zeros_like(x)
	res[:, 0] += other_info['karman']['v_magnitude']
	return res


def load_vortex_particles(path):
	"""`v x _ y w` records of the particle list (2D/init_cond.py:213-224): positions (M,2) and strengths (M)"""
	X, Y, W = [], [], []
	with open(path, 'r') as fd:
		for line in fd.readlines():
			if line.startswith('v '):
				t = line.split(' ')
				X.append(float(t[1])); Y.append(float(t[3])); W.append(float(t[4]))
	return torch.tensor([X, Y]).transpose(0, 1).contiguous(), torch.tensor(W)


def vortices_pass_particles_of(pos, strength):
	"""2D/init_cond.py:226-238: u(x) = rot90( sum_j w_j (p_j - x) / (|p_j - x|^2 + eps) ), eps = .1; the Jacobian in closed form
	(the reference differentiates with torch.func.jacfwd)"""
	eps = .1

	def field(x, grad):
		p, w = pos.to(x.device), strength.to(x.device)
		d = p[None] - x[:, None]	# (Q, M, 2)
		q = (d ** 2).sum(dim=-1) + eps
		if not grad:
			s = (w[None, :, None] * d / q[..., None]).sum(dim=1)
			return torch.stack([-s[:, 1], s[:, 0]], dim=1).contiguous()
		# ds_k/dx_l = sum_j w_j ( -delta_kl / q + 2 d_k d_l / q^2 )
		ds = (w[None, :, None, None] * (-torch.eye(2, device=x.device)[None, None] / q[..., None, None] + 2. * d[..., :, None] * d[..., None, :] / (q ** 2)[..., None, None])).sum(dim=1)
		return torch.stack([-ds[:, 1, :], ds[:, 0, :]], dim=1).contiguous()
	return field


FIELDS = {'taylor_green': taylor_green, 'taylor_vortex': taylor_vortex, 'leapfrog': leapfrog, 'karman': karman,
		  'vortices_pass': vortices_pass_of('vortices_pass'), 'vortices_pass_narrow': vortices_pass_of('vortices_pass_narrow'),
		  'vortices_pass_noslip': vortices_pass_of('vortices_pass_noslip')}


class Scene2D:
	"""everything 2D/init_cond.py derives from `--init_cond`"""

	def __init__(self, init_cond, particles_obj=None):
		if init_cond not in initialize_domain:
			raise KeyError(f'unknown 2D scene {init_cond!r}')
		self.name = init_cond
		self.info = other_info[init_cond]
		self.initialize_domain = initialize_domain[init_cond]
		self.advance_domain = list(advance_domain[init_cond])	# karman moves its left edge (extra_advector)
		self.visualize_domain = visualize_domain[init_cond]
		self.particle_count = initial_particle_count[init_cond]
		self.visualize_res = visualize_res[init_cond]
		x_min, x_max, y_min, y_max = self.initialize_domain
		self.scaling_factor = 10. / min(x_max - x_min, y_max - y_min)	# 2D/init_cond.py:22-25
		if init_cond == 'vortices_pass_particles':
			self._field = vortices_pass_particles_of(*load_vortex_particles(particles_obj or self.info['particles_obj']))
		else:
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

	target_velocity.graph_safe = target_gradient.graph_safe = True

	def scaled(self, dom):
		return tuple(v * self.scaling_factor for v in dom)

	# ---- the moving inlet of karman (2D/init_cond.py:267-300) ------------------------------------------------------
	def extra_advector(self, dt, advection_scheme='rk4'):
		if self.name == 'karman':
			self.advance_domain[0] = min(self.advance_domain[0] + dt * self.info['v_magnitude'], self.visualize_domain[0])

	def extra_loader(self, start_frame, dt):
		if self.name == 'karman':
			self.advance_domain[0] = min(self.initialize_domain[0] + (start_frame * dt) * self.info['v_magnitude'], self.visualize_domain[0])

	# ---- samplers -------------------------------------------------------------------------------------------------
	def data_generator(self, gaussian_splatting, domain=None):
		"""default_data_generator of 2D/advance.py:314-316: Q = N uniform samples of the advance domain, GSR space (the initial fit,
		2D/initialize.py:216-217, draws them on the initialize domain: pass it as `domain`)"""
		x_min, x_max, y_min, y_max = domain if domain is not None else self.advance_domain
		dev = _dev()
		sf = self.scaling_factor	# (rand * extent + lo) * sf, with the constants folded: two kernels
		return torch.addcmul(_const([x_min * sf, y_min * sf], dev), torch.rand_like(gaussian_splatting.positions.detach(), device=dev),
							 _const([(x_max - x_min) * sf, (y_max - y_min) * sf], dev))

	def test_generator(self):
		x_min, x_max, y_min, y_max = self.advance_domain
		return gsr2d.get_grid_points(x_min, x_max, y_min, y_max, *self.visualize_res) * self.scaling_factor

	def _on_domain_boundary_2(self, n):
		"""sample_on_domain_boundary_2 (2D/init_cond.py:306-325): points on the four edges (perimeter-weighted), OUTWARD normals,
		target normal velocity 0.  One torch.rand like the reference; the edge of a draw is a table lookup (corner + direction x
		arc length) instead of the reference's four masked assignments: a dozen small kernels, no boolean-mask indexing (no host sync)."""
		x_min, x_max, y_min, y_max = self.advance_domain
		xs, ys = x_max - x_min, y_max - y_min
		dev = _dev()
		t = torch.rand(n, device=dev) * ((xs + ys) * 2.)
		edge = torch.bucketize(t, _const([xs, xs + ys, 2. * xs + ys], dev), right=True)	# t >= bound -> next edge
		s_ = t - _const([0., xs, xs + ys, 2. * xs + ys], dev)[edge]
		corner = _const([[x_min, y_min], [x_max, y_min], [x_max, y_max], [x_min, y_max]], dev)[edge]
		along = _const([[1., 0.], [0., 1.], [-1., 0.], [0., -1.]], dev)[edge]
		data = torch.addcmul(corner, along, s_[:, None])
		normals = _const([[0., -1.], [1., 0.], [0., 1.], [-1., 0.]], dev)[edge]
		return data, normals, torch.zeros(n, device=dev)

	@staticmethod
	def _on_circle(n, x, y, r):
		"""the points of sample_on_sphere_1 / _2 (2D/init_cond.py:327-343): returns (data, unit outward normal)"""
		dev = _dev()
		theta = torch.rand(n, device=dev) * 2. * np.pi
		nrm = torch.stack([torch.cos(theta), torch.sin(theta)], dim=1)
		return r * nrm + _const([x, y], dev), nrm

	def _discs(self):
		return [self.info[k] for k in ('obstacle_pos1', 'obstacle_pos2') if k in self.info] or [self.info['obstacle_pos']]

	def _discs_1(self, n):
		"""value samples on the discs: u = 0 there (sample_for_vortices_pass_1 / sample_for_karman_1, :345-352, :391-392)"""
		data = torch.cat([self._on_circle(n, x, y, self.info['obstacle_radius'])[0] for (x, y) in self._discs()], dim=0)
		return data, torch.zeros_like(data)

	def _discs_2(self, n):
		"""normal samples on the discs: u.n = 0 (the disc parts of sample_for_vortices_pass_2 / _particles_2, :354-372)"""
		parts = [self._on_circle(n, x, y, self.info['obstacle_radius']) for (x, y) in self._discs()]
		return torch.cat([p[0] for p in parts], dim=0), torch.cat([p[1] for p in parts], dim=0), torch.zeros(n * len(parts), device=_dev())

	def _karman_2(self, n):
		"""sample_for_karman_2 (2D/init_cond.py:394-425): channel walls (u.n = 0), inlet and outlet of the advance domain and the
		left edge of the visualised window (u.n = +-v_magnitude)"""
		x_min, x_max, y_min, y_max = self.advance_domain
		x_min_v, vm = self.visualize_domain[0], self.info['v_magnitude']
		dev = _dev()
		t = torch.rand(n, device=dev) * (x_max - x_min) + x_min
		t2 = torch.rand(n, device=dev) * (y_max - y_min) + y_min
		col = lambda v: torch.full((n,), v, device=dev)
		data = torch.cat([torch.stack([t, col(y_min)], 1), torch.stack([t, col(y_max)], 1), torch.stack([col(x_min), t2], 1),
						  torch.stack([col(x_max), t2], 1), torch.stack([col(x_min_v), t2], 1)], dim=0)
		nrm = _const([[0., 1.], [0., -1.], [1., 0.], [-1., 0.], [1., 0.]], dev).repeat_interleave(n, dim=0)
		val = _const([0., 0., vm, -vm, vm], dev).repeat_interleave(n)
		return data, nrm, val

	def _raw_samplers(self):
		"""[boundary_generator_1, boundary_generator_2] in original coordinates (the `boundary_sampler` table, :440-449)"""
		def pass_2(n):	# discs first, then the domain box: the reference's order of draws (:354-362)
			d, nr, v = self._discs_2(n)
			d3, n3, v3 = self._on_domain_boundary_2(n)
			return torch.cat([d, d3], dim=0), torch.cat([nr, n3], dim=0), torch.cat([v, v3], dim=0)
		return {'vortices_pass': [None, pass_2], 'vortices_pass_narrow': [None, pass_2],
				'vortices_pass_noslip': [self._discs_1, self._on_domain_boundary_2],
				'vortices_pass_particles': [None, self._discs_2],
				'karman': [self._discs_1, self._karman_2]}.get(self.name, [None, self._on_domain_boundary_2])

	def boundary_sampler_2(self, n):
		"""the scene's normal-velocity sampler in GSR space (target_boundary_sampler_2, :433-438): (points, normals, target u.n)"""
		data, normal, val = self._raw_samplers()[1](n)
		return (data * self.scaling_factor).contiguous(), normal.contiguous(), val * self.scaling_factor

	def boundary_sampler_1(self, n):
		"""the scene's value sampler in GSR space (target_boundary_sampler_1, :427-432): (points, target u), or None for free-slip scenes"""
		raw = self._raw_samplers()[0]
		if raw is None:
			return None
		data, value = raw(n)
		return (data * self.scaling_factor).contiguous(), (value * self.scaling_factor).contiguous()

	# pure device functions of torch's CUDA random stream: the optimisation loops may replay them from a CUDA graph (advance2d.project)
	boundary_sampler_1.graph_safe = boundary_sampler_2.graph_safe = True

	@property
	def boundary_samplers(self):
		"""[boundary_generator_1, boundary_generator_2] as in 2D/init_cond.py:440-449"""
		return [self.boundary_sampler_1 if self._raw_samplers()[0] is not None else None, self.boundary_sampler_2]
