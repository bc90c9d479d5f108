"""
3D initial fit — the roles of the reference's 3D/initialize.py: fit_velocity_with_gradient (:9-46) and SimulationInitialize
(:49-100, without the VTI dumps of the analytic field, which are output only).

fused=True runs each epoch as forward -> atomics-free backward (value L1 + gradient L1, the `weight_val` / `weight_grad`
branches of 3D/GSR.py:396-411, :436-451) -> gsr_step_rebuild (closed-form anisotropy / volume regularisers, Adam x4,
ReduceLROnPlateau, hash rebuild) with no host synchronisation; fused=False is the reference's formulation on the drop-in
class (torch autograd regularisers, torch.optim.Adam) and the parity partner in the tests.
"""
import time

import torch
import torch.nn.functional as F

from . import gsr3d, init_cond3d
from .engine import FusedStepper
from .graphloop import GraphedLoop


def fit_velocity_with_gradient(gaussian_velocity, reference_field, reference_gradient, data_generator, batch_size=8192, max_epoch=3000, verbose=1, fused=True,
							   use_graph=None):
	"""the initial fit (3D/initialize.py:9-46).  fused=True: one device-resident iteration (forward, gather, gsr_step_rebuild) without host
	synchronisation; use_graph (None: when the generator and the target field carry `graph_safe = True`): ten iterations per captured CUDA
	graph, the next batch and its analytic targets prepared on a second stream (graphloop.py) — same results bit for bit"""
	gv = gaussian_velocity
	gv.initialize_optimizers()
	dev = gsr3d.device
	if not fused:
		for epoch in range(max_epoch):
			data = data_generator(batch_size)
			ref_val, ref_grad = reference_field(data), reference_gradient(data)
			val, grad = gv.get_losses(data, ref_val=ref_val, weight_val=1., ref_grad=ref_grad, weight_grad=1.)
			ratio = torch.exp(gv.scalings.max(dim=-1).values - gv.scalings.min(dim=-1).values)
			volumes = torch.exp(-gv.scalings.sum(dim=-1))
			loss_aniso = (torch.where(ratio >= 1.5, ratio, torch.full_like(ratio, 1.5)) - 1.5).mean()
			loss_vol = ((volumes / volumes.mean() - 1) ** 2).mean()
			(loss_aniso + loss_vol).backward()
			gv.step(F.l1_loss(val, ref_val) + F.l1_loss(grad, ref_grad) + loss_aniso + loss_vol)
		return
	e = gv._engine
	stepper = FusedStepper(e, [gv.positions_lr, gv.scalings_lr, gv.rotations_lr, gv.values_lr], 50, 1., 1., pcgrad=False,
						   tau=gv.clamp_threshold, min_grid_scale=gv.min_grid_scale, ext_bounds=gv._ext())
	stepper.init(gv.scalings)
	e.build(gv.positions.detach(), params=[p.detach() for p in gv._params()])
	e._packed_key = None
	st_time = time.time()
	def batches():	# the samples and the analytic targets at them: functions of the random stream only
		data = data_generator(batch_size).detach()
		if hasattr(reference_field, 'both'):	# the analytic ring fields give velocity and Jacobian from one pass over the particles
			ref_val, ref_grad = reference_field.both(data)
		else:
			ref_val, ref_grad = reference_field(data).contiguous(), reference_gradient(data).contiguous()
		return data, ref_val, ref_grad

	def iteration(inputs):
		data, ref_val, ref_grad = inputs
		Q = data.shape[0]
		bins = e.bin_samples(data, True)
		if 'val' not in bufs:
			bufs['val'], bufs['grad'] = torch.empty((Q, 3), device=dev), torch.empty((Q, 3, 3), device=dev)
		val, grad = bufs['val'], bufs['grad']
		e.forward(data, val, grad, accumulate=False, perm=bins)
		acc, mask = e.backward_gather(data, bins.perm, bins.scs, val, grad, (1., 0., 1., 0., 0., 0.), {'ref_val': ref_val, 'ref_grad': ref_grad}, None, want_losses=True)
		lp, nblk = e.last_loss_partials
		stepper.step([p.detach() for p in gv._params()], acc, mask, loss_srcs=[(lp, nblk, [0., 0., 0., 0., 1. / Q, 1. / Q, 0., 0.])], rebuild=True)
	bufs = {}
	if use_graph is None:
		use_graph = getattr(data_generator, 'graph_safe', False) and getattr(reference_field, 'graph_safe', False)
	loop = GraphedLoop(iteration, unit=10, enabled=bool(use_graph), prepare=batches)
	done = 0
	while done < max_epoch:
		k = min(100, max_epoch - done)
		loop.run(k)
		done += k
		if verbose:
			sc = stepper.scalars()
			print(f'loss_tot: {sc[9]}, loss_aniso: {sc[10]}, loss_vol: {sc[11]}, time: {time.time() - st_time}')
			st_time = time.time()
	loop.release()
	gv.grid_scale = stepper.detach()
	e._packed_key = None
	for p in gv._params():
		p.add_(0.)	# modified through raw pointers: bump the autograd version counters
	gv.zero_grad()


def simulation_initialize(init_cond, max_epoch=500, verbose=1, fused=True, particle_count=None, use_graph=None):
	"""lattice of Gaussians over the scene's domain -> fit to the analytic vortex-ring field -> the frame-0 field (3D/initialize.py:49-86)"""
	x_min, x_max, y_min, y_max, z_min, z_max = init_cond3d.domain[init_cond]
	nx, ny, nz = particle_count or init_cond3d.initial_particle_count[init_cond]
	field = init_cond3d.make_field(init_cond)
	pts = gsr3d.get_grid_points(x_min, x_max, y_min, y_max, z_min, z_max, nx, ny, nz).cpu().numpy()
	gv = gsr3d.GaussianSplatting3DFast(x_min, x_max, y_min, y_max, z_min, z_max, pts, dim=3)
	dev = gsr3d.device
	ext = torch.tensor([x_max - x_min, y_max - y_min, z_max - z_min], device=dev)
	lo = torch.tensor([x_min, y_min, z_min], device=dev)
	gen = lambda n: torch.rand_like(gv.positions.detach(), device=dev) * ext + lo	# Q = N: batch_size is ignored, as in the reference (:73-74)
	gen.graph_safe = True
	fit_velocity_with_gradient(gv, field, field.gradient, gen, max_epoch=max_epoch, verbose=verbose, fused=fused, use_graph=use_graph)
	return gv
