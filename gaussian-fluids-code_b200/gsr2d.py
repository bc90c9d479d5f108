"""
Drop-in replacement of the reference's 2D/GSR.py (angle-parametrised Gaussians, separate value / gradient passes) on the
same sm_100a CUDA engine as gsr3d.  Same deliberate differences as gsr3d (no argv parsing at import, dim == 2 only for
the Fast class, deterministic sums); the plotting helpers of the reference (show_field, draw_ellipses) are out of scope.
"""
import argparse
import os

import numpy as np
import torch

from . import host
from .engine import HashEngine
from ._lib import GsrError


def parse_args(argv=None):
	"""same flags and defaults as 2D/GSR.py:13-21"""
	parser = argparse.ArgumentParser()
	parser.add_argument('--device', type=str, default='0')
	parser.add_argument('--dir', type=str, default='output_fast')
	parser.add_argument('--start_frame', type=int, default=0)
	parser.add_argument('--init_cond', type=str, default='taylor_vortex')
	parser.add_argument('--dt', type=float, default=.01)
	parser.add_argument('--last_time', type=float, default=10.)
	return parser.parse_args(argv)


cmd_args = parse_args([])
device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')


def configure(argv=None, make_dir=True, seed=42):
	"""what the reference does at import time (2D/GSR.py:22-30)"""
	global cmd_args, device
	args = parse_args(argv)
	cmd_args.__dict__.update(args.__dict__)
	if make_dir:
		os.makedirs(cmd_args.dir, exist_ok=True)
	torch.manual_seed(seed)
	if cmd_args.device != 'cpu':
		if not torch.cuda.is_initialized():
			os.environ['CUDA_VISIBLE_DEVICES'] = cmd_args.device
		torch.cuda.manual_seed_all(seed)
	device = torch.device('cpu' if cmd_args.device == 'cpu' else 'cuda')
	return cmd_args


class GaussianSplatting:
	"""dense representation + optimiser plumbing — the role of 2D/GSR.py:35-169"""

	def __init__(self, positions, dim):
		self.N, self.dim = positions.shape[0], dim
		self.positions = torch.tensor(positions, dtype=torch.float, requires_grad=True, device=device)
		self.scalings = torch.zeros((self.N, 2), requires_grad=True, device=device)	# log INVERSE radii
		self.rotations = torch.zeros(self.N, requires_grad=True, device=device)
		self.values = torch.zeros((self.N, dim), requires_grad=True, device=device)

	def set_lr(self, positions_lr, scalings_lr, rotations_lr, values_lr):
		self.positions_lr, self.scalings_lr, self.rotations_lr, self.values_lr = positions_lr, scalings_lr, rotations_lr, values_lr

	def initialize_optimizers(self, patience=50):
		self.optimizers, self.schedulers = [], []
		for name in ('positions', 'scalings', 'rotations', 'values'):
			opt = torch.optim.Adam([getattr(self, name)], lr=getattr(self, name + '_lr'))
			sch = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=.9, patience=patience)
			setattr(self, name + '_optimizer', opt)
			setattr(self, name + '_scheduler', sch)
			self.optimizers.append(opt)
			self.schedulers.append(sch)

	def parameters(self):
		return {'positions': self.positions, 'scalings': self.scalings, 'rotations': self.rotations, 'values': self.values}

	def save(self, filename):
		torch.save(self.parameters(), filename)

	def load(self, filename):
		d = torch.load(filename, map_location=device)
		self.positions, self.scalings, self.rotations, self.values = d['positions'], d['scalings'], d['rotations'], d['values']
		self.N, self.dim = self.positions.shape[0], self.values.shape[1]

	def get_scaling_matrices(self):
		return torch.diag_embed(torch.exp(self.scalings))

	def get_rotation_matrices(self):
		c, s = torch.cos(self.rotations), torch.sin(self.rotations)
		return torch.stack((c, -s, s, c), dim=-1).reshape(-1, 2, 2)

	def get_variances(self):
		"""Sigma^-1 = (R S)(R S)^T (the reference's name for it)"""
		A = self.get_rotation_matrices() @ self.get_scaling_matrices()
		return A @ A.transpose(-1, -2)

	def _dense_terms(self, x):
		d = x[:, None, :] - self.positions[None, :, :]
		w = torch.einsum('nkl,qnl->qnk', self.get_variances(), d)
		g = torch.exp(-.5 * (d * w).sum(-1))
		return self.values[None] * g[..., None], w

	def forward_single(self, x):
		"""u at ONE point x (2,) -> (dim,)  (2D/GSR.py:110-113)"""
		return self._dense_terms(x[None])[0].sum(dim=1)[0]

	def __call__(self, x):
		if x.dim() == 1:
			return self.forward_single(x)
		return self._dense_terms(x)[0].sum(dim=1)

	def gradient_single(self, x, need_val=False):
		"""grad u at ONE point x (2,) -> (dim, 2) [, u]  (2D/GSR.py:123-132)"""
		return GaussianSplatting.gradient(self, x, need_val)

	def gradient(self, x, need_val=False):
		single = x.dim() == 1
		per, w = self._dense_terms(x[None] if single else x)
		grad, val = -(per[..., :, None] * w[..., None, :]).sum(dim=1), per.sum(dim=1)
		if single:
			grad, val = grad[0], val[0]
		return (grad, val) if need_val else grad

	def freeze(self):
		for p in GaussianSplatting.parameters(self).values():
			p.requires_grad_(False)

	def unfreeze(self):
		for p in GaussianSplatting.parameters(self).values():
			p.requires_grad_()

	def zero_grad(self):
		for o in self.optimizers:
			o.zero_grad()

	def step(self, metrics):
		for o in self.optimizers:
			o.step()
		for s in self.schedulers:
			s.step(metrics)


class GaussianSplattingFast(GaussianSplatting):
	"""truncated, hashed 2D representation on the CUDA engine — the role of 2D/GSR.py:171-647"""

	def __init__(self, x_min, x_max, y_min, y_max, positions, min_grid_scale=None, clamp_threshold=1e-3, dim=1, load_file=None):
		super().__init__(positions, dim)
		self._engine = HashEngine(2, device)
		if load_file is None:
			bounds = (x_min, x_max, y_min, y_max)
			self.min_grid_scale = host.default_min_grid_scale(2, bounds, self.N) if min_grid_scale is None else min_grid_scale
			self.clamp_threshold = clamp_threshold
			self.x_min, self.x_max, self.y_min, self.y_max = host.extend(2, bounds, self.min_grid_scale)
			with torch.no_grad():
				self.scalings += host.initial_scaling(self.clamp_threshold, self.min_grid_scale)
			self.create_grid_data()
			self.zero_grad()
		else:
			self.load(load_file)

	def _ext(self):
		return [self.x_min, self.x_max, self.y_min, self.y_max]

	def create_grid_data(self):
		self.grid_size = host.grid_size(2, self._ext(), self.min_grid_scale)

	def reinitialize_grid(self):
		"""2D/GSR.py:224-229"""
		if self.dim != 2:
			raise GsrError('GaussianSplattingFast supports dim == 2 only')
		min_s = self._engine.min_scaling(self.scalings.detach()).item() if self.clamp_threshold else 0.
		self.grid_scale = host.grid_scale(self.clamp_threshold, min_s, self.min_grid_scale, self._ext())
		self._engine.set_grid(self._ext(), self.grid_size, self.grid_scale, self.clamp_threshold)
		self._engine.build(self.positions.detach())

	def grid_arrays(self):
		cnt, off = self._engine.build(self.positions.detach(), want_ref_format=True)
		total = int(self._engine.cell_start[-1].item())
		return cnt.reshape(self.grid_size), off.reshape(self.grid_size), self._engine.sorted_id[:total].clone()

	def _params(self):
		return (self.positions.detach(), self.scalings.detach(), self.rotations.detach(), self.values.detach())

	def parameters(self):
		d = super().parameters()
		d.update({'clamp_threshold': self.clamp_threshold, 'min_grid_scale': self.min_grid_scale,
				  'domain_range': (self.x_min, self.x_max, self.y_min, self.y_max)})
		return d

	def load(self, filename):
		d = torch.load(filename, map_location=device)
		self.positions, self.scalings, self.rotations, self.values = d['positions'], d['scalings'], d['rotations'], d['values']
		self.N, self.dim = self.positions.shape[0], self.values.shape[1]
		self.clamp_threshold, self.min_grid_scale = d['clamp_threshold'], d['min_grid_scale']
		self.x_min, self.x_max, self.y_min, self.y_max = d['domain_range']
		self.create_grid_data()
		self.zero_grad()

	def _grads(self):
		return [self.positions.grad, self.scalings.grad, self.rotations.grad, self.values.grad]

	def get_losses(self, x, ref=None, weight=0., normals=None, normal_ref=None, weight_boundary=0., stop_gradient=None):
		"""value pass: u, then the value-L1 and boundary |u.n - n_ref| backward (2D/GSR.py:341-360)"""
		if ref is None:
			weight = 0.
		if normals is None or normal_ref is None:
			weight_boundary = 0.
		e = self._engine
		e.ensure_packed(self._params())
		backward = weight != 0. or weight_boundary != 0.
		val = torch.zeros((x.shape[0], self.dim), device=device)
		bins = e.bin_samples(x.detach(), need_cells=backward)
		perm, scs = bins
		e.forward(x, val, None, accumulate=False, perm=bins)
		if backward:
			refs = {'ref_val': ref if weight != 0. else None, 'normals': normals if weight_boundary != 0. else None,
					'normal_ref': normal_ref if weight_boundary != 0. else None}
			acc, mask = e.backward_gather(x, perm, scs, val, None, (weight, weight_boundary, 0., 0., 0., 0.), refs, stop_gradient)
			e.backward_epilogue(self.scalings, self.rotations, acc, mask, [self._grads(), None, None])
		return val

	def __call__(self, x):
		return self.get_losses(x)

	def get_coverage(self, x):
		"""sum_i (g_i(x) - tau)_+ at every sample: how much of the representation covers x (2D/GSR.py:594-618).  The forward kernel
		on records packed with v = (1, 1): both components of its output are the coverage"""
		e = self._engine
		ones = torch.ones((self.N, 2), device=device)
		e.ensure_packed([self.positions.detach(), self.scalings.detach(), self.rotations.detach(), ones])
		out = torch.zeros((x.shape[0], 2), device=device)
		e.forward(x.detach().contiguous(), out, None, accumulate=False, perm=e.bin_samples(x.detach().contiguous(), need_cells=False))
		e._packed_key = None	# the records hold the unit values: re-pack before the next evaluation of the field itself
		return out[:, 0].contiguous()

	def get_grad_losses(self, x, ref_grad=None, weight_grad=0., ref_vor=None, weight_vor=0., weight_div=0.,
						vor_positions_grad=None, vor_scalings_grad=None, vor_rotations_grad=None, vor_values_grad=None,
						div_positions_grad=None, div_scalings_grad=None, div_rotations_grad=None, div_values_grad=None, stop_gradient=None):
		"""gradient pass: grad u, then the gradient-L1, vorticity-L1 and divergence-L2 backward (2D/GSR.py:478-521)"""
		if ref_grad is None:
			weight_grad = 0.
		if ref_vor is None:
			weight_vor = 0.
		e = self._engine
		e.ensure_packed(self._params())
		backward = weight_grad != 0. or weight_vor != 0. or weight_div != 0.
		grad = torch.zeros((x.shape[0], self.dim, 2), device=device)
		bins = e.bin_samples(x.detach(), need_cells=backward)
		perm, scs = bins
		e.forward(x, None, grad, accumulate=False, perm=bins)
		if backward:
			refs = {'ref_grad': ref_grad if weight_grad != 0. else None, 'ref_vor': ref_vor if weight_vor != 0. else None}
			acc, mask = e.backward_gather(x, perm, scs, None, grad, (0., 0., weight_grad, weight_vor, 0., weight_div), refs, stop_gradient)
			direct = self._grads()
			pick = lambda given, k: given if given is not None else direct[k]
			vor = [pick(vor_positions_grad, 0), pick(vor_scalings_grad, 1), pick(vor_rotations_grad, 2), pick(vor_values_grad, 3)]
			div = [pick(div_positions_grad, 0), pick(div_scalings_grad, 1), pick(div_rotations_grad, 2), pick(div_values_grad, 3)]
			e.backward_epilogue(self.scalings, self.rotations, acc, mask, [direct, vor, div])
		return grad

	def gradient(self, x, need_val=False):
		grad = self.get_grad_losses(x)
		return (grad, self.__call__(x)) if need_val else grad

	def advection_rk4(self, start_pos, dt, pos_only=True):
		"""2D/GSR.py:582-592"""
		e = self._engine
		e.ensure_packed(self._params())
		Q = start_pos.shape[0]
		goal_pos = torch.zeros_like(start_pos, device=device)
		if pos_only:
			e.rk4(start_pos, dt, goal_pos)
			return goal_pos
		deformation = torch.zeros((Q, 2, 2), device=device)
		goal_val = torch.zeros((Q, 2), device=device)
		goal_grad = torch.zeros((Q, 2, 2), device=device)
		e.rk4(start_pos, dt, goal_pos, deformation, goal_val, goal_grad)
		return goal_pos, deformation, goal_val, goal_grad

	def advected_vorticity(self, x, dt, domain=None):
		"""fused AdvectedCovectorField.vorticity (2D/advance.py:46-54): curl at the point back-traced by -dt, zero outside `domain`"""
		e = self._engine
		e.ensure_packed(self._params())
		vor = torch.empty((x.shape[0],), device=device)
		e.advected_vorticity(x, -dt, vor, None, domain=domain)
		return vor

	def get_all_neighbors(self, x):
		"""2D/GSR.py:632-635 (returns bool like the reference)"""
		e = self._engine
		e.ensure_packed(self._params())
		mark = torch.zeros((self.positions.shape[0],), dtype=torch.int32, device=device)
		e.mark_neighbors(x, mark)
		return mark.bool()

	def zero_grad(self):
		"""2D/GSR.py:637-643"""
		for param in GaussianSplatting.parameters(self).values():
			if param.grad is None:
				param.grad = torch.zeros_like(param, device=device)
			else:
				param.grad.zero_()
		self.reinitialize_grid()

	def step(self, metrics):
		super().step(metrics)
		self.zero_grad()


def get_grid_points(x_min, x_max, y_min, y_max, x_N, y_N):
	"""lattice of x_N*y_N points, x fastest: point i*x_N + j = (X[j], Y[i]) (2D/GSR.py:667-672, meshgrid indexing 'xy')"""
	X = torch.linspace(x_min, x_max, x_N, device=device)
	Y = torch.linspace(y_min, y_max, y_N, device=device)
	Yg, Xg = torch.meshgrid(Y, X, indexing='ij')
	return torch.stack((Xg, Yg), dim=-1).reshape(-1, 2).contiguous()
