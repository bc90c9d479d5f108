"""
Drop-in replacement of the reference's 3D/GSR.py: same classes, methods, arguments and return values, with the
Taichi kernels replaced by the sm_100a CUDA engine behind include/gsr_b200.h.

Differences that are deliberate:
  * importing this module does not parse sys.argv or create directories (the reference does both at import,
    3D/GSR.py:12-30); `cmd_args` holds the same defaults and `configure()` parses a command line on request;
  * only `dim == 3` is supported by the Fast class (every caller of the reference uses dim == 3);
  * the intra-cell order of sorted_id is canonical (ascending id) and all sums are deterministic.
"""
import argparse
import os

import numpy as np
import torch

from . import host
from .engine import HashEngine
from ._lib import GsrError


def parse_args(argv=None):
	"""same flags and defaults as 3D/GSR.py:12-21"""
	parser = argparse.ArgumentParser()
	parser.add_argument('--device', type=str, default='0')
	parser.add_argument('--dir', type=str, default='output_3d')
	parser.add_argument('--start_frame', type=int, default=0)
	parser.add_argument('--boundary', type=float, default=10.)
	parser.add_argument('--init_cond', type=str, default='leapfrog')
	parser.add_argument('--dt', type=float, default=.02)
	parser.add_argument('--last_time', type=float, default=100.)
	return parser.parse_args(argv)


cmd_args = parse_args([])
device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')


def configure(argv=None, make_dir=True, seed=42):
	"""What the reference does at import time (3D/GSR.py:22-30): parse flags, make --dir, seed, pick the device."""
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


class GaussianSplatting3D:
	"""Dense (untruncated) representation and the optimiser plumbing — the role of 3D/GSR.py:34-152."""

	def __init__(self, positions, dim=1, positions_lr=1.6e-3, scalings_lr=5e-2, rotations_lr=5e-2, values_lr=5e-3):
		self.N, self.dim = positions.shape[0], dim
		self.positions = torch.tensor(positions, device=device, requires_grad=True)
		self.scalings = torch.zeros((self.N, 3), device=device, requires_grad=True)
		rot = torch.zeros((self.N, 4), device=device)
		rot[:, 0] = 1.
		self.rotations = rot.requires_grad_()
		self.values = torch.zeros((self.N, dim), device=device, requires_grad=True)
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
		q = self.rotations / self.rotations.norm(dim=-1, keepdim=True)
		r, x, y, z = q.unbind(-1)
		rows = [1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
				2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
				2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]
		return torch.stack(rows, dim=-1).reshape(-1, 3, 3)

	def get_variances(self):
		"""returns Sigma^-1 = (R S)(R S)^T despite the name — callers use it as a precision matrix (3D/advance.py:69)"""
		A = self.get_rotation_matrices() @ self.get_scaling_matrices()
		return A @ A.transpose(-1, -2)

	def _dense_terms(self, x):
		d = x[:, None, :] - self.positions[None, :, :]
		w = torch.einsum('nkl,qnl->qnk', self.get_variances(), d)
		g = torch.exp(-.5 * (d * w).sum(-1))
		return self.values[None] * g[..., None], w

	def __call__(self, x):
		return self._dense_terms(x)[0].sum(dim=1)

	def gradient(self, x, need_val=False):
		per, w = self._dense_terms(x)
		grad = -(per[..., :, None] * w[..., None, :]).sum(dim=1)
		return (grad, per.sum(dim=1)) if need_val else grad

	def freeze(self):
		for p in GaussianSplatting3D.parameters(self).values():	# the Fast subclass adds scalars to parameters()
			p.requires_grad_(False)

	def unfreeze(self):
		for p in GaussianSplatting3D.parameters(self).values():
			p.requires_grad_()

	def zero_grad(self):
		for o in self.optimizers:
			o.zero_grad()

	def step(self, metrics):
		for o in self.optimizers:
			o.step()
		for s in self.schedulers:
			s.step(metrics)


class GaussianSplatting3DFast(GaussianSplatting3D):
	"""Truncated, spatially hashed representation on the CUDA engine — the role of 3D/GSR.py:154-716."""

	def __init__(self, x_min, x_max, y_min, y_max, z_min, z_max, positions, min_grid_scale=None, clamp_threshold=5e-3, dim=1,
				 positions_lr=1e-3, scalings_lr=1e-3, rotations_lr=1e-3, values_lr=1e-3, load_file=None):
		super().__init__(positions, dim, positions_lr, scalings_lr, rotations_lr, values_lr)
		self._engine = HashEngine(3, device)
		if load_file is None:
			bounds = (x_min, x_max, y_min, y_max, z_min, z_max)
			self.min_grid_scale = host.default_min_grid_scale(3, bounds, self.N) if min_grid_scale is None else min_grid_scale
			self.clamp_threshold = clamp_threshold
			self.x_min, self.x_max, self.y_min, self.y_max, self.z_min, self.z_max = host.extend(3, bounds, self.min_grid_scale)
			with torch.no_grad():
				self.scalings += host.initial_scaling(self.clamp_threshold, self.min_grid_scale)
			self.create_grid_data()
			self.zero_grad()
		else:
			self.load(load_file)

	# ---- hash -------------------------------------------------------------------------------------
	def _ext(self):
		return [self.x_min, self.x_max, self.y_min, self.y_max, self.z_min, self.z_max]

	def create_grid_data(self):
		self.grid_size = host.grid_size(3, self._ext(), self.min_grid_scale)

	def reinitialize_grid(self):
		"""3D/GSR.py:247-252 — the scalar is formed in host double precision exactly as the reference does"""
		if self.dim != 3:
			raise GsrError('GaussianSplatting3DFast supports dim == 3 only')
		min_s = self._engine.min_scaling(self.scalings.detach()).item() if self.clamp_threshold else 0.
		self.grid_scale = host.grid_scale(self.clamp_threshold, min_s, self.min_grid_scale, self._ext())
		self._engine.set_grid(self._ext(), self.grid_size, self.grid_scale, self.clamp_threshold)
		self._engine.build(self.positions.detach())

	def grid_arrays(self):
		"""(grid_cnt, grid_offset, sorted_id) in the reference's format, for inspection and parity tests"""
		cnt, off = self._engine.build(self.positions.detach(), want_ref_format=True)
		total = int(self._engine.cell_start[-1].item())
		return cnt.reshape(self.grid_size), off.reshape(self.grid_size), self._engine.sorted_id[:total].clone()

	def _params(self):
		return (self.positions.detach(), self.scalings.detach(), self.rotations.detach(), self.values.detach())

	def parameters(self):
		d = super().parameters()
		d.update({'clamp_threshold': self.clamp_threshold, 'min_grid_scale': self.min_grid_scale,
				  'domain_range': (self.x_min, self.x_max, self.y_min, self.y_max, self.z_min, self.z_max)})
		return d

	def load(self, filename, first_time=True):
		d = torch.load(filename, map_location=device)
		self.positions, self.scalings, self.rotations, self.values = d['positions'], d['scalings'], d['rotations'], d['values']
		self.N, self.dim = self.positions.shape[0], self.values.shape[1]
		self.clamp_threshold, self.min_grid_scale = d['clamp_threshold'], d['min_grid_scale']
		self.x_min, self.x_max, self.y_min, self.y_max, self.z_min, self.z_max = d['domain_range']
		if first_time:
			self.create_grid_data()
		self.zero_grad()

	# ---- kernels ----------------------------------------------------------------------------------
	def get_losses(self, x, discard_grad=False,
				   ref_val=None, weight_val=0., normals=None, weight_boundary=0., ref_grad=None, weight_grad=0.,
				   ref_vor=None, weight_vor=0., ref_hel=None, weight_hel=0., weight_div=0.,
				   vor_positions_grad=None, vor_scalings_grad=None, vor_rotations_grad=None, vor_values_grad=None,
				   div_positions_grad=None, div_scalings_grad=None, div_rotations_grad=None, div_values_grad=None,
				   stop_gradient=None):
		"""3D/GSR.py:542-597: forward u (and grad u), then the analytic backward of the weighted losses"""
		e = self._engine
		e.ensure_packed(self._params())
		Q = x.shape[0]
		val = torch.zeros((Q, self.dim), device=device)
		grad = torch.zeros((0 if discard_grad else Q, self.dim, 3), device=device)
		weights = (weight_val, weight_boundary, weight_grad, weight_vor, weight_hel, weight_div)
		# 3D/GSR.py:299 — weight_hel alone does not enable the backward loop
		backward = any(w != 0. for w in (weight_val, weight_boundary, weight_grad, weight_vor, weight_div))
		if backward and discard_grad:
			raise GsrError('a backward pass needs grad (discard_grad=False)')
		bins = e.bin_samples(x.detach(), need_cells=backward)
		perm, scs = bins
		e.forward(x, val, None if discard_grad else grad, accumulate=True, perm=bins)
		if backward:
			refs = {'ref_val': ref_val if weight_val != 0. else None, 'normals': normals if weight_boundary != 0. else None,
					'ref_grad': ref_grad if weight_grad != 0. else None, 'ref_vor': ref_vor if weight_vor != 0. else None,
					'ref_hel': ref_hel if weight_hel != 0. else None}
			acc, mask = e.backward_gather(x, perm, scs, val, grad, weights, refs, stop_gradient)
			direct = [self.positions.grad, self.scalings.grad, self.rotations.grad, self.values.grad]
			pick = lambda given, k: given if given is not None else direct[k]
			vor = [pick(vor_positions_grad, 0), pick(vor_scalings_grad, 1), pick(vor_rotations_grad, 2), pick(vor_values_grad, 3)]
			div = [pick(div_positions_grad, 0), pick(div_scalings_grad, 1), pick(div_rotations_grad, 2), pick(div_values_grad, 3)]
			e.backward_epilogue(self.scalings, self.rotations, acc, mask, [direct, vor, div])
		return val if discard_grad else (val, grad)

	def advection_rk4(self, start_pos, dt, pos_only=True):
		"""3D/GSR.py:667-677"""
		e = self._engine
		e.ensure_packed(self._params())
		Q = start_pos.shape[0]
		goal_pos = torch.zeros_like(start_pos, device=device)
		if pos_only:
			e.rk4(start_pos, dt, goal_pos)
			return goal_pos
		deformation = torch.zeros((Q, 3, 3), device=device)
		goal_val = torch.zeros((Q, 3), device=device)
		goal_grad = torch.zeros((Q, 3, 3), device=device)
		e.rk4(start_pos, dt, goal_pos, deformation, goal_val, goal_grad)
		return goal_pos, deformation, goal_val, goal_grad

	def advected_vorticity(self, x, dt, need_hel=False):
		"""fused form of AdvectedCovectorField.vorticity (3D/advance.py:24-47): back-trace by -dt, pull back the curl"""
		e = self._engine
		e.ensure_packed(self._params())
		Q = x.shape[0]
		vor = torch.empty((Q, 3), device=device)
		hel = torch.empty((Q,), device=device) if need_hel else None
		e.advected_vorticity(x, -dt, vor, hel)
		return (vor, hel) if need_hel else vor

	def get_all_neighbors(self, x):
		"""3D/GSR.py:692-695 (returns int32 like the reference)"""
		e = self._engine
		e.ensure_packed(self._params())
		mark = torch.zeros((self.positions.shape[0],), dtype=torch.int32, device=device)
		e.mark_neighbors(x, mark)
		return mark

	def __call__(self, x):
		return self.get_losses(x, discard_grad=True)

	def gradient(self, x, need_val=False):
		val, grad = self.get_losses(x)
		return (grad, val) if need_val else grad

	def zero_grad(self):
		"""3D/GSR.py:704-712: lazily create the optimisers, zero (or create) .grad, rebuild the hash"""
		if not hasattr(self, 'optimizers'):
			self.initialize_optimizers()
		for param in super().parameters().values():
			if param.grad is None:
				param.grad = torch.zeros_like(param, device=device)
			else:
				param.grad.zero_()
		self.reinitialize_grid()

	def step(self, metrics):
		super().step(metrics)
		self.zero_grad()


def get_grid_points(x_min, x_max, y_min, y_max, z_min, z_max, x_N, y_N, z_N):
	"""lattice of x_N*y_N*z_N points, z fastest (3D/GSR.py:719-725)"""
	axes = [torch.linspace(a, b, n, device=device) for a, b, n in ((x_min, x_max, x_N), (y_min, y_max, y_N), (z_min, z_max, z_N))]
	return torch.stack(torch.meshgrid(*axes, indexing='ij'), dim=-1).reshape(-1, 3).contiguous()


def write_vti(field, x_min, x_max, y_min, y_max, z_min, z_max, save_filename, x_N=30, y_N=30, z_N=30):
	"""
	Samples `field` on the lattice and writes a VTK ImageData file (3D/GSR.py:728-742).  VTK is not a dependency here:
	the XML is written by hand (ascii, point data, x fastest as VTK expects).
	"""
	XYZ = get_grid_points(x_min, x_max, y_min, y_max, z_min, z_max, x_N, y_N, z_N)
	write_vti_array(field(XYZ).reshape(x_N, y_N, z_N), x_min, x_max, y_min, y_max, z_min, z_max, save_filename)


def write_vti_array(values, x_min, x_max, y_min, y_max, z_min, z_max, save_filename):
	"""the file of write_vti for values already sampled on the (x_N, y_N, z_N) lattice of get_grid_points"""
	x_N, y_N, z_N = values.shape
	V = values.detach().cpu().numpy()
	sx, sy, sz = (x_max - x_min) / x_N, (y_max - y_min) / y_N, (z_max - z_min) / z_N
	data = ' '.join(f'{v:.7g}' for v in V.ravel(order='F'))
	with open(save_filename, 'w') as fd:
		fd.write('<?xml version="1.0"?>\n<VTKFile type="ImageData" version="0.1" byte_order="LittleEndian">\n')
		fd.write(f'<ImageData WholeExtent="0 {x_N - 1} 0 {y_N - 1} 0 {z_N - 1}" Origin="{x_min} {y_min} {z_min}" Spacing="{sx} {sy} {sz}">\n')
		fd.write(f'<Piece Extent="0 {x_N - 1} 0 {y_N - 1} 0 {z_N - 1}">\n<PointData Scalars="scalars">\n')
		fd.write(f'<DataArray type="Float32" Name="scalars" format="ascii">\n{data}\n</DataArray>\n</PointData>\n</Piece>\n</ImageData>\n</VTKFile>\n')


def write_obj(gs, save_filename):
	P = gs.positions.detach().cpu().numpy()
	with open(save_filename, 'w') as fd:
		for p in P:
			fd.write(f'v {p[0]} {p[1]} {p[2]}\n')
