"""
Passive density advection — the role of the reference's 3D/advance_density.py: the smoke rings of
`ring_collide` are carried by the saved velocity fields, frame after frame, on a lattice 4x finer than the visualisation
lattice (512^3 by default).

The reference materialises the lattice and the back-traced lattice (2 x 1.6 GB) and resamples with a Taichi kernel; here
`gsr_advect_density` does the RK4 back-trace (advection_rk4 with -dt, 3D/advance_density.py:54), the clamp (:55) and the
trilinear resampling (ti_get_interp_val, :25-50) in one kernel per step, for both rings at once.
"""
import os

import numpy as np
import torch

from . import gsr3d, init_cond3d


class DensityAdvector:
	def __init__(self, x_min, x_max, y_min, y_max, z_min, z_max, res=(512, 512, 512)):
		self.domain = (x_min, x_max, y_min, y_max, z_min, z_max)
		self.res = tuple(res)
		dev = gsr3d.device
		# the axis values of get_grid_points (torch.linspace, 3D/GSR.py:719-725) — the kernel reads the coordinates from these
		self.axes = tuple(torch.linspace(lo, hi, n, device=dev) for lo, hi, n in ((x_min, x_max, res[0]), (y_min, y_max, res[1]), (z_min, z_max, res[2])))

	def set_ring(self, ring):
		"""ti_set_ring (3D/advance_density.py:13-22): 1 inside the torus of the given centre line radius and tube thickness"""
		dev = gsr3d.device
		c = torch.tensor(ring['center'], dtype=torch.float32, device=dev)
		nrm = torch.tensor(ring['normal'], dtype=torch.float32, device=dev)
		radius, thick = float(ring['radius']), float(ring['thickness'])
		xs, ys, zs = self.axes
		out = torch.zeros(self.res, device=dev)
		for i0 in range(0, self.res[0], 32):	# slabs: bounded temporaries at 512^3
			X = torch.stack(torch.meshgrid(xs[i0:i0 + 32], ys, zs, indexing='ij'), dim=-1)
			proj = X - ((X - c) * nrm).sum(-1, keepdim=True) * nrm
			rad = proj - c
			rl = rad.norm(dim=-1, keepdim=True)
			nearest = c + rad / rl * radius
			out[i0:i0 + 32] = ((rl[..., 0] >= radius - thick) & ((X - nearest).norm(dim=-1) <= thick)).float()
		return out

	def advect(self, gaussian_velocity, dt, density_a, density_b=None, x_range=None, out=None):
		"""advected_density (3D/advance_density.py:52-58) for one or two fields; returns the new field(s).  x_range = (begin, end):
		only those x planes are computed — the rest of the returned field(s) is left as `out` holds it (SlabSharding)"""
		gv = gaussian_velocity
		gv._engine.ensure_packed(gv._params())
		out_a = out[0] if out is not None else torch.empty_like(density_a)
		out_b = (out[1] if out is not None else torch.empty_like(density_b)) if density_b is not None else None
		gv._engine.advect_density(self.axes, self.domain, -dt, density_a, out_a, density_b, out_b, x_range=x_range)
		return out_a if density_b is None else (out_a, out_b)


class SlabSharding:
	"""
	The density lattice shared between the GPUs of a box by slabs of x planes (config 5 of BASELINE.json: 512^3 on 8 GPUs): every
	rank advects its own slab; a back-traced voxel reads the old density at most `halo` planes outside the slab (the flow moves
	|u| dt per frame, a fraction of a plane at the reference's settings), so after every frame the ranks exchange `halo` planes
	with their two neighbours — point-to-point, no collective.  Every rank keeps full-shape arrays of which its slab +- halo is valid.
	"""

	def __init__(self, nx, rank, world, halo=4):
		self.nx, self.rank, self.world, self.halo = nx, rank, world, halo
		self.begin, self.end = rank * nx // world, (rank + 1) * nx // world

	def x_range(self):
		return self.begin, self.end

	def exchange(self, *fields):
		"""send this slab's outermost `halo` planes to the neighbours, receive theirs (torch.distributed point-to-point)"""
		if self.world == 1:
			return
		import torch.distributed as dist
		h, ops = self.halo, []
		for f in fields:
			if self.rank > 0:
				ops.append(dist.P2POp(dist.isend, f[self.begin:self.begin + h], self.rank - 1))
				ops.append(dist.P2POp(dist.irecv, f[self.begin - h:self.begin], self.rank - 1))
			if self.rank < self.world - 1:
				ops.append(dist.P2POp(dist.isend, f[self.end - h:self.end], self.rank + 1))
				ops.append(dist.P2POp(dist.irecv, f[self.end:self.end + h], self.rank + 1))
		for r in dist.batch_isend_irecv(ops):
			r.wait()


def advected_density_reference(density, gaussian_velocity, dt, domain):
	"""the reference's own composition (materialised lattice, advection_rk4, clamp, trilinear taps written with torch) — the
	parity partner of gsr_advect_density in the tests; O(lattice) temporaries, small lattices only"""
	x_min, x_max, y_min, y_max, z_min, z_max = domain
	nx, ny, nz = density.shape
	dev = density.device
	x = gsr3d.get_grid_points(x_min, x_max, y_min, y_max, z_min, z_max, nx, ny, nz)
	bk = gaussian_velocity.advection_rk4(x, -dt)
	lo = torch.tensor([x_min, y_min, z_min], device=dev)
	hi = torch.tensor([x_max, y_max, z_max], device=dev)
	bk = torch.minimum(torch.maximum(bk, lo), hi)
	d = (hi - lo) / torch.tensor([nx - 1, ny - 1, nz - 1], dtype=torch.float32, device=dev)
	p = bk - lo
	idx = torch.floor(p / d).long()
	n1 = torch.tensor([nx - 1, ny - 1, nz - 1], device=dev)
	idx = torch.minimum(idx.clamp_min(0), n1)
	idx1 = torch.minimum(idx + 1, n1)
	w = (p - d * idx.float()) / d
	res = torch.zeros(x.shape[0], device=dev)
	for a, ia, wa in ((0, idx, 1. - w), (1, idx1, w)):
		for b, ib, wb in ((0, idx, 1. - w), (1, idx1, w)):
			for c_, ic, wc in ((0, idx, 1. - w), (1, idx1, w)):
				res += density[ia[:, 0], ib[:, 1], ic[:, 2]] * wa[:, 0] * wb[:, 1] * wc[:, 2]
	return res.reshape(nx, ny, nz)


def tensor2vti(V, x_min, x_max, y_min, y_max, z_min, z_max, save_filename):
	"""3D/advance_density.py:74-85 without VTK: binary-free ImageData XML (ascii), point data in x-fastest order"""
	nx, ny, nz = V.shape
	data = ' '.join(f'{v:.6g}' for v in V.detach().cpu().numpy().ravel(order='F'))
	with open(save_filename, 'w') as fd:
		fd.write('<?xml version="1.0"?>\n<VTKFile type="ImageData" version="0.1" byte_order="LittleEndian">\n')
		fd.write(f'<ImageData WholeExtent="0 {nx - 1} 0 {ny - 1} 0 {nz - 1}" Origin="{x_min} {y_min} {z_min}" '
				 f'Spacing="{(x_max - x_min) / nx} {(y_max - y_min) / ny} {(z_max - z_min) / nz}">\n')
		fd.write(f'<Piece Extent="0 {nx - 1} 0 {ny - 1} 0 {nz - 1}">\n<PointData Scalars="scalars">\n')
		fd.write(f'<DataArray type="Float32" Name="scalars" format="ascii">\n{data}\n</DataArray>\n</PointData>\n</Piece>\n</ImageData>\n</VTKFile>\n')


def run(directory, init_cond='ring_collide', dt=.02, res=None, write=True):
	"""the `__main__` loop of 3D/advance_density.py:87-119: advect the two rings through gaussian_velocity_{0,1,...}.pt"""
	if init_cond != 'ring_collide':
		raise NotImplementedError	# as in the reference (:88-96)
	dom = init_cond3d.domain[init_cond]
	res = res or tuple(4 * n for n in init_cond3d.visualize_res[init_cond])
	adv = DensityAdvector(*dom, res=res)
	info = init_cond3d.other_info[init_cond]
	d1, d2 = adv.set_ring(info['ring1']), adv.set_ring(info['ring2'])
	frame = 0
	gv = None
	while os.path.exists(os.path.join(directory, f'gaussian_velocity_{frame}.pt')):
		fn = os.path.join(directory, f'gaussian_velocity_{frame}.pt')
		if gv is None:
			gv = gsr3d.GaussianSplatting3DFast(*dom, np.zeros((1, 3), np.float32), dim=3, load_file=fn)
		else:
			gv.load(fn)
		frame += 1
		d1, d2 = adv.advect(gv, dt, d1, d2)
		if write:
			tensor2vti(d1, *dom, os.path.join(directory, f'density_a_{frame}.vti'))
			tensor2vti(d2, *dom, os.path.join(directory, f'density_b_{frame}.vti'))
	return d1, d2, frame
