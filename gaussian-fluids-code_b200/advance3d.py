"""
Drop-in replacement of the reference's 3D/advance.py: AdvectedCovectorField, clone_velocity_field,
advect_covector_field and project keep their names, arguments and behaviour.

project() has three executions of the same iteration:
  * pipelined (fused=True and the STOCK generators of this module — BoxSampler / LatticeGenerator / BoxSurfaceSampler, the
    objects advance() builds, standing for default_data_generator / default_test_generator / sample_on_box of
    3D/advance.py:339-342 and 3D/init_cond.py:227-249): the samples are drawn on the device by counter-based Philox kernels,
    ten iterations form one captured CUDA graph with the pull-back, the forward pass and the boundary pass on forked streams
    (timestep3d.ShardedProjector), the host only replays the graph and reads three numbers every `check_iter` iterations for
    the early-stop rule.  This is the measured path of bench.py.
  * fused=True with arbitrary generator callables: the same device-resident iteration (sample binning, RK4 pull-back of the
    previous field, forward, atomics-free backward, boundary pass, then ONE fused step — PCGrad projection, closed-form
    regularisers, Adam x4, ReduceLROnPlateau x4, next grid_scale — and the hash rebuild), launched eagerly because the
    generators run on the host side every iteration.
  * fused=False: the reference's structure (get_losses + torch autograd regularisers + torch.optim) on top of the
    same CUDA kernels — >= 8 host syncs per iteration like the reference; kept for API fidelity and as a cross-check.

advance() / advance_frame() are the time loop of 3D/advance.py:381-393 as functions (the reference runs it at module level).

Deliberate deviations from the reference, both in clone_velocity_field: `test_data` is
generated before its first use, and the neighbour mask is converted to bool before `~` (the 2D reference does both).
"""
import os
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import gsr3d
from . import _lib
from .engine import FusedStepper
from .gsr3d import GaussianSplatting3DFast, get_grid_points  # noqa: F401  (re-exported like the reference's star import)


def _dev():
	return gsr3d.device


def curl(jacob):
	"""omega_k from grad u[j, d, l] = d u_d / d x_l"""
	return torch.stack((jacob[:, 2, 1] - jacob[:, 1, 2], jacob[:, 0, 2] - jacob[:, 2, 0], jacob[:, 1, 0] - jacob[:, 0, 1]), dim=-1)


class AdvectedCovectorField:
	"""origin_covector_field advected by velocity_field for time_step seconds (3D/advance.py:11-49)"""

	def __init__(self, origin_covector_field, velocity_field, time_step, x_min, x_max, y_min, y_max, z_min, z_max, advection_scheme='rk4'):
		self.origin_covector_field = origin_covector_field
		self.velocity_field = velocity_field
		self.time_step = time_step
		self.x_min, self.x_max, self.y_min, self.y_max, self.z_min, self.z_max = x_min, x_max, y_min, y_max, z_min, z_max
		self.advection_scheme = advection_scheme

	def vorticity(self, x, need_hel=False):
		"""vorticity (and helicity) of the advected covector field at x; x is left untouched"""
		if self.advection_scheme != 'rk4':
			raise NotImplementedError
		# RK4 back-trace by -dt, curl of the pulled-back Jacobian, Dpsi^-1 and u.curl fused in one kernel
		return self.velocity_field.advected_vorticity(x, self.time_step, need_hel=need_hel)



# ---- the stock generators of the reference's driver, as objects that project() recognises ----------------------------------------
class BoxSampler:
	"""default_data_generator of 3D/advance.py:339-340: one uniform sample per Gaussian in the box (`n` is ignored there too)"""

	def __init__(self, x_min, x_max, y_min, y_max, z_min, z_max):
		self.box = (float(x_min), float(x_max), float(y_min), float(y_max), float(z_min), float(z_max))
		self.gsr_sampler = ('box', self.box)

	def __call__(self, n, gaussian_splatting, restrict=None):
		b, dev = self.box, _dev()
		return (torch.rand_like(gaussian_splatting.positions.detach(), device=dev) * torch.tensor([b[1] - b[0], b[3] - b[2], b[5] - b[4]], device=dev)
				+ torch.tensor([b[0], b[2], b[4]], device=dev))


class LatticeGenerator:
	"""default_test_generator of 3D/advance.py:341-342: the visualisation lattice (the same tensor on every call, so the engine
	keeps its ordering); `points` may hold this process's share of the lattice in a sharded run"""

	fixed_points = True	# every call returns the same points: project() evaluates the pull-back reference on them once per phase

	def __init__(self, x_min, x_max, y_min, y_max, z_min, z_max, x_N, y_N, z_N, points=None, total=None):
		self.args = (x_min, x_max, y_min, y_max, z_min, z_max, x_N, y_N, z_N)
		self._points = points
		self.total = int(total if total is not None else x_N * y_N * z_N)	# points of the whole lattice (all processes)

	def __call__(self, gaussian_splatting=None):
		if self._points is None:
			self._points = get_grid_points(*self.args)
		return self._points


class BoxSurfaceSampler:
	"""sample_on_box (3D/init_cond.py:227-249) on a fixed box: n points on the faces with inward normals"""

	def __init__(self, x_min, x_max, y_min, y_max, z_min, z_max):
		self.box = (float(x_min), float(x_max), float(y_min), float(y_max), float(z_min), float(z_max))
		self.gsr_sampler = ('box_surface', self.box)

	def __call__(self, n):
		from .init_cond3d import sample_on_box
		return sample_on_box(n, *self.box)


def _stock(gen, kind):
	tag = getattr(gen, 'gsr_sampler', None)
	return tag is not None and tag[0] == kind


def _regularisers(scalings, stop_gradient=None):
	"""anisotropy hinge at ratio 1.5 and volume uniformity (3D/advance.py:107-115, :237-241)"""
	aniso_ratio = 1.5
	s = scalings if stop_gradient is None else scalings[~stop_gradient]
	ratio = torch.exp(s.max(dim=-1).values - s.min(dim=-1).values)
	if not ratio.shape[0]:
		ratio = torch.ones((1,), device=_dev())
	loss_aniso = (torch.where(ratio >= aniso_ratio, ratio, aniso_ratio) - aniso_ratio).mean()
	if stop_gradient is None:
		volumes = torch.exp(-scalings.sum(dim=-1))
	else:
		volumes = torch.where(stop_gradient, torch.exp(-scalings.detach().sum(dim=-1)), torch.exp(-scalings.sum(dim=-1)))
	loss_vol = ((volumes / volumes.mean() - 1) ** 2).mean()
	return loss_aniso, loss_vol


def clone_velocity_field(res, velocity_field, x_min, x_max, y_min, y_max, z_min, z_max, data_generator, test_data_generator,
						 reinitialize=False, batch_size=8192, max_epoch=3000, patience=500, verbose=1, normals=None, seed=0):
	"""
	Copy velocity_field into res, split over-stretched Gaussians (axis ratio >= 2, repeatedly) into two samples of their own
	distribution and fit the new ones — and their neighbours — to the old field with the value + gradient losses while everything
	else stays frozen (3D/advance.py:51-165).  The split runs on the device (reseed.py / csrc/split.cu); `normals` (a list of
	(2, n_split, 3) standard-normal draws, one per round) and `seed` are for reproducing a split.
	"""
	from . import reseed
	names = ('positions', 'scalings', 'rotations', 'values')
	with torch.no_grad():
		same = res is not velocity_field and all(getattr(res, nm).shape == getattr(velocity_field, nm).shape and getattr(res, nm).requires_grad for nm in names)
		if same:
			# copy INTO res's tensors: the values are what the reference's `.clone()` gives, and the storage stays where a captured
			# iteration graph of this field expects it (a split below replaces the tensors, as in the reference)
			for nm in names:
				getattr(res, nm).copy_(getattr(velocity_field, nm))
		else:
			for nm in names:
				setattr(res, nm, getattr(velocity_field, nm).detach().clone())
		res.N = res.positions.shape[0]
		stop_gradient, n_split = reseed.split_all(res, 3, clamp_box=(res.x_min, res.x_max, res.y_min, res.y_max, res.z_min, res.z_max), normals=normals, seed=seed,
												  verbose=verbose)
	res.unfreeze()
	res.zero_grad()
	if n_split == 0:
		return res
	# the neighbours of the new Gaussians train too (3D/advance.py:101; the reference applies `~` to an int32 mask there: B.3)
	stop_gradient = torch.logical_and(stop_gradient, ~res.get_all_neighbors(res.positions[~stop_gradient].detach().contiguous()).bool())

	def losses(data, backward):
		ref_val, ref_grad = velocity_field.get_losses(data)
		if backward:
			val, grad = res.get_losses(data, ref_val=ref_val, weight_val=1., ref_grad=ref_grad, weight_grad=1., stop_gradient=stop_gradient)
		else:
			val, grad = res.get_losses(data)
		loss_val, loss_grad = F.l1_loss(val, ref_val), F.l1_loss(grad, ref_grad)
		loss_aniso, loss_vol = _regularisers(res.scalings, stop_gradient)
		if backward:
			(loss_aniso + loss_vol).backward()
		return loss_val + loss_grad + loss_aniso + loss_vol, loss_val, loss_grad, loss_aniso, loss_vol

	res.positions_lr = res.rotations_lr = res.scalings_lr = res.values_lr = 1e-3	# 3D/advance.py:118-121
	res.initialize_optimizers()
	for s in res.schedulers:
		s.factor = .9
	reseed.refit(res, losses, data_generator, test_data_generator, ~stop_gradient, batch_size, max_epoch, patience, verbose)
	return res


def advect_covector_field(covector_field, velocity_field, dt, x_min=None, x_max=None, y_min=None, y_max=None, z_min=None, z_max=None, advection_scheme='rk4'):
	"""move the Gaussians of covector_field along velocity_field (RK4, positions only), clamp, rebuild the hash (3D/advance.py:167-180)"""
	if advection_scheme != 'rk4':
		raise NotImplementedError
	new_positions = velocity_field.advection_rk4(covector_field.positions.detach(), dt)
	if x_min is not None:
		device = _dev()
		new_positions.clamp_(torch.tensor([x_min, y_min, z_min], dtype=torch.float32, device=device), torch.tensor([x_max, y_max, z_max], dtype=torch.float32, device=device))
	if covector_field.positions.requires_grad and covector_field.positions.shape == new_positions.shape:
		with torch.no_grad():
			covector_field.positions.copy_(new_positions)	# same values as the reference's rebinding, same storage for captured graphs
	else:
		new_positions.requires_grad_()
		covector_field.positions = new_positions
	covector_field.zero_grad()


# the pull-back reference of a FIXED test lattice is evaluated once per project phase instead of at every test pass (FusedProjector.evaluate)
HOIST_TEST_REFERENCE = os.environ.get('GSR_HOIST_TEST_REFERENCE', '1') != '0'
PROJECT_WEIGHTS = dict(vor=1., hel=1., div=1., aniso=10., vol=10., val_reg=0.)	# 3D/advance.py:184
PROJECT_LRS = dict(positions=3e-4, scalings=1e-5, rotations=3e-4, values=1e-5)	# 3D/advance.py:258-261


class FusedProjector:
	"""One device-resident `project` phase for gaussian_velocity against the advected previous field."""

	def __init__(self, gaussian_velocity, reference_field, boundary_lambda=0., patience=50, weights=None, lrs=None):
		gv = self.gv = gaussian_velocity
		self.ref = reference_field
		self.w = dict(PROJECT_WEIGHTS, **(weights or {}))
		self.lrs = dict(PROJECT_LRS, **(lrs or {}))
		self.boundary_lambda = float(boundary_lambda)
		e = gv._engine
		self.stepper = FusedStepper(e, [self.lrs[k] for k in ('positions', 'scalings', 'rotations', 'values')], patience,
									self.w['aniso'], self.w['vol'], w_valreg=self.w['val_reg'], pcgrad=True,
									tau=gv.clamp_threshold, min_grid_scale=gv.min_grid_scale, ext_bounds=gv._ext(), sample_grid_ahead=True)
		self.stepper.init(gv.scalings)
		self._it = 0	# iterations since the phase began (host count; inside replayed graphs only its parity is meaningful)
		self._rebuild()
		cur = reference_field.velocity_field
		cur._engine.ensure_packed(cur._params())
		self._buf = {}
		self._streams = None

	def sample_grid(self, iteration=None):
		"""The grid scale the sample batches of an iteration are binned with (a device scalar): not the hash's own but one the step
		kernel left an iteration EARLIER (gsr_step_cfg.sample_gs_slots: slot [iteration & 1] = (1 + margin) x the previous grid_scale,
		checked against the next one) — so a batch can be drawn and ordered before the step that rebuilds the hash has finished
		(timestep3d.ShardedProjector does).  Every form of the iteration uses it, which keeps them bit-identical."""
		it = self._it if iteration is None else iteration
		return self.stepper.sample_gs[(it & 1):(it & 1) + 1]

	def _rebuild(self):
		gv = self.gv
		e = gv._engine
		e.build(gv.positions.detach(), params=[p.detach() for p in gv._params()])	# hash + packed records, one call
		e._packed_key = None	# parameters are updated in place by raw pointers: never trust the version counters here

	def _tmp(self, name, shape):
		t = self._buf.get(name)
		if t is None or tuple(t.shape) != tuple(shape):
			t = torch.empty(shape, dtype=torch.float32, device=_dev())
			self._buf[name] = t
		return t

	def iterate(self, data, boundary=None):
		"""one optimiser iteration; no host synchronisation.  The boundary pass, the forward pass of the training batch and the RK4
		pull-back reference are independent until the training gather / the step: three streams, joined by events (a captured graph keeps
		the concurrency; same kernels on the same inputs as the sequential order, same bits)"""
		gv, e = self.gv, self.gv._engine
		cur = self.ref.velocity_field
		Q = data.shape[0]
		data = data.detach()
		sgs = self.sample_grid()
		self._it += 1
		main = torch.cuda.current_stream()
		if self._streams is None:
			self._streams = (torch.cuda.Stream(), torch.cuda.Stream())
		# (eagerly the host is the bottleneck: forking streams only adds event calls — 375 instead of 200 ms per S1 frame — so the
		# chains are forked only while the iteration is being captured, where the graph then runs them side by side)
		multi = torch.cuda.is_current_stream_capturing()
		s_fwd, s_bnd = self._streams if multi else (main, main)
		fork = binned = done_f = None
		if multi:
			fork = torch.cuda.Event()
			fork.record(main)
		extra, srcs_b, done_b = [], [], None
		if boundary is not None and self.boundary_lambda:
			bdata, bnormal = boundary
			bdata, bnormal = bdata.detach(), bnormal.detach()
			Qb = bdata.shape[0]
			if multi:
				s_bnd.wait_event(fork)
			with torch.cuda.stream(s_bnd):
				bins_b = e.bin_samples(bdata, True, tag='b', gs_dev=sgs)
				perm_b, scs_b = bins_b
				valb = self._tmp('valb', (Qb, 3))
				e.forward(bdata, valb, None, accumulate=False, perm=bins_b)
				acc_b, mask_b = e.backward_gather(bdata, perm_b, scs_b, valb, None, (0., self.boundary_lambda, 0., 0., 0., 0.),
												  {'normals': bnormal}, None, tag='acc_b', want_losses=True, sample_gs=sgs)
				lpb, nblkb = e.last_loss_partials
				if multi:
					done_b = torch.cuda.Event()
					done_b.record(s_bnd)
			srcs_b.append((lpb, nblkb, [0., 0., 0., self.boundary_lambda / Qb, 0., 0., 0., 0.]))
			extra.append(acc_b)
		bins = e.bin_samples(data, True, gs_dev=sgs)
		perm, scs = bins
		ref_vor, ref_hel = self._tmp('ref_vor', (Q, 3)), self._tmp('ref_hel', (Q,))
		val, grad = self._tmp('val', (Q, 3)), self._tmp('grad', (Q, 3, 3))
		if multi:
			binned = torch.cuda.Event()
			binned.record(main)
			s_fwd.wait_event(binned)
		with torch.cuda.stream(s_fwd):
			e.forward(data, val, grad, accumulate=False, perm=bins)
			if multi:
				done_f = torch.cuda.Event()
				done_f.record(s_fwd)
		cur._engine.advected_vorticity(data, -self.ref.time_step, ref_vor, ref_hel, perm=bins)	# beside the forward pass
		if multi:
			main.wait_event(done_f)
		acc, mask = e.backward_gather(data, perm, scs, val, grad, (0., 0., 0., self.w['vor'], self.w['hel'], self.w['div']),
									  {'ref_vor': ref_vor, 'ref_hel': ref_hel}, None, want_losses=True, sample_gs=sgs)
		lp, nblk = e.last_loss_partials
		srcs = [(lp, nblk, [self.w['vor'] / Q, 0., self.w['div'] / Q, 0., 0., 0., 0., 0.])] + srcs_b
		if done_b is not None:
			main.wait_event(done_b)
		self.stepper.step([p.detach() for p in gv._params()], acc, mask, extra=extra, loss_srcs=srcs, rebuild=True)	# update + hash + packed records

	def evaluate(self, data, probe=None, fixed=False):
		"""losses of the current field on `data` without a gradient (the test pass, 3D/advance.py:226-235): device sums.
		probe: optional list that receives a (start, end) CUDA-event pair around the RK4 pull-back kernel (bench.py).
		fixed: the caller vouches that `data` holds the same points as at the previous call of this phase (LatticeGenerator).  The
		pull-back reference omega(phi(x)) reads only those points and the PREVIOUS field, so it is then computed by the first test
		pass of the phase and reused by the later ones (the reference recomputes the identical values at every pass); a new phase
		(restart) always recomputes it.  HOIST_TEST_REFERENCE = False restores the recomputation."""
		gv, e = self.gv, self.gv._engine
		data = data.detach()
		Q = data.shape[0]
		bins = e.bin_samples(data, False)
		perm = bins.perm
		ref_vor, ref_hel = self._tmp('t_ref_vor', (Q, 3)), self._tmp('t_ref_hel', (Q,))
		key = (data.data_ptr(), Q, id(self.ref), float(self.ref.time_step), ref_vor.data_ptr()) if (fixed and HOIST_TEST_REFERENCE) else None
		self.reference_reused = key is not None and key == getattr(self, '_test_ref_key', None)
		if not self.reference_reused:
			if probe is not None:
				ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
				ev[0].record()
			self.ref.velocity_field._engine.advected_vorticity(data, -self.ref.time_step, ref_vor, ref_hel, perm=bins)
			if probe is not None:
				ev[1].record()
				probe.append(ev)
			self._test_ref_key = key
		val, grad = self._tmp('t_val', (Q, 3)), self._tmp('t_grad', (Q, 3, 3))
		e.forward(data, val, grad, accumulate=False, perm=bins)
		return e.sample_losses(val, grad, {'ref_vor': ref_vor, 'ref_hel': ref_hel}, Q) / Q

	def check_sample_grid(self):
		"""raises if a step found the hash's grid_scale above the sample grid a batch had been ordered with (synchronises)"""
		if float(self.stepper.state[_lib.ST_SGS_ERR].item()) != 0.:
			raise _lib.GsrError('grid_scale grew by more than the sample grid margin within one optimiser step: the learning rate of the '
								'scalings is too large for exp(4 lr) to bound it')

	def finish(self):
		"""return control to the generic API: host grid_scale, fresh hash, version-tracked packing"""
		self.check_sample_grid()
		gv = self.gv
		gv.grid_scale = self.stepper.detach()
		gv._engine._packed_key = None
		for p in gv._params():
			p.add_(0.)	# bump the autograd version counters: the tensors were modified through raw pointers


def project(gaussian_velocity, reference_field, x_min, x_max, y_min, y_max, z_min, z_max, data_generator, test_data_generator,
			boundary_generator=None, boundary_lambda=0., batch_size=8192, max_epoch=3000, patience=500, verbose=1, frame_id=None,
			fused=True, check_iter=100, history=None, pipelined=None, rank=0, world=1, lattice_world=None, sample_seed=42, census=None, probe=None,
			use_graph=True):
	"""
	Solve one time step's projection by first-order optimisation (3D/advance.py:182-316): fit the vorticity and helicity
	of the advected covector field while driving the divergence to zero.  Early stop: every `check_iter` iterations the
	test losses must improve by 0.1 %, or `patience` iterations without improvement on all three end the phase.
	Returns the number of iterations run.  (The reference's loss-curve PNG, :317-331, is not produced; pass a dict as
	`history` to collect the same series.)

	Beyond the reference's signature: `pipelined` (None = automatically when the generators are this module's stock objects)
	selects the captured-graph execution; rank / world shard the samples over processes (the losses are means over samples, so the
	normalisers use the global counts and ONE exchange per iteration sums the gradient accumulators: timestep3d); lattice_world
	(default: world) is the number of processes that share the test lattice — the test generator then returns this process's share
	and carries the whole lattice's point count in `.total`; sample_seed keys the device sample generators; census / probe /
	use_graph are measurement hooks of bench.py.
	"""
	gv = gaussian_velocity
	gv.positions_lr, gv.scalings_lr, gv.rotations_lr, gv.values_lr = [PROJECT_LRS[k] for k in ('positions', 'scalings', 'rotations', 'values')]
	gv.initialize_optimizers(patience=50)
	for s in gv.schedulers:
		s.factor = .9
	if not fused:
		return _project_unfused(gv, reference_field, data_generator, test_data_generator, boundary_generator, boundary_lambda,
								batch_size, max_epoch, patience, verbose, check_iter, history)
	use_boundary = bool(boundary_lambda and boundary_generator)
	stock = _stock(data_generator, 'box') and (not use_boundary or _stock(boundary_generator, 'box_surface'))
	if pipelined is None:
		pipelined = stock
	if pipelined:
		if not stock:
			raise ValueError('the pipelined projection draws its samples on the device: pass BoxSampler / BoxSurfaceSampler generators')
		return _project_pipelined(gv, reference_field, data_generator, test_data_generator, boundary_generator if use_boundary else None, boundary_lambda,
								  batch_size, max_epoch, patience, verbose, check_iter, history, rank, world, world if lattice_world is None else lattice_world,
								  sample_seed, census, probe, use_graph)
	if world > 1:
		raise ValueError('sharded projection needs the pipelined path')
	fp = FusedProjector(gv, reference_field, boundary_lambda if boundary_generator else 0., patience=50)
	names = ('loss_vor', 'loss_hel', 'loss_div')
	fixed_test = bool(getattr(test_data_generator, 'fixed_points', False))	# a LatticeGenerator: the pull-back reference is evaluated once
	if verbose:
		t = fp.evaluate(test_data_generator(gv), fixed=fixed_test).tolist()
		print(f'[projection] loss_vor: {t[0]}, loss_hel: {t[1]}, loss_div: {t[2]}')
	best = {k: np.inf for k in names}
	stale = {k: 0 for k in names}
	st_time = time.time()
	epochs = max_epoch
	use_b = bool(boundary_lambda and boundary_generator)

	def batches():
		return data_generator(batch_size, gv), (boundary_generator(batch_size) if use_b else None)

	def iteration(inputs):
		fp.iterate(*inputs)
	# generators that are pure device functions of the CUDA random stream (`graph_safe = True`: init_cond3d.make_boundary_sampler's are)
	# let the loop replay from CUDA graphs, ten iterations each (whole pairs: the sample grids alternate with the iteration's parity)
	from .graphloop import GraphedLoop
	safe = getattr(data_generator, 'graph_safe', False) and (not use_b or getattr(boundary_generator, 'graph_safe', False))
	unit = next((u for u in (10, 4, 2) if check_iter % u == 0), 0)
	loop = GraphedLoop(iteration, unit=unit or 1, enabled=bool(use_graph and safe and unit), prepare=batches)
	done = 0
	while done < max_epoch:
		k = min(check_iter, max_epoch - done)
		loop.run(k)
		done += k
		if k < check_iter:
			break
		t = fp.evaluate(test_data_generator(gv), fixed=fixed_test).tolist()	# the only host sync of the loop
		cur = dict(zip(names, t[:3]))
		if history is not None:
			history.setdefault('test', []).append(cur)
			history.setdefault('state', []).append(fp.stepper.scalars()[:16])
		if verbose:
			print(f'[projection] loss_vor: {t[0]}, loss_hel: {t[1]}, loss_div: {t[2]}, time: {time.time() - st_time}')
			st_time = time.time()
		for kk in names:
			if cur[kk] < best[kk] * (1. - 1e-3):
				best[kk], stale[kk] = cur[kk], 0
			else:
				stale[kk] += check_iter
		if all(stale[kk] >= patience for kk in names):
			epochs = done
			if verbose:
				print('[projection] Total epoch:', epochs)
			break
	else:
		if verbose:
			print('[projection] Total epoch:', max_epoch, '(Reached maximum iteration number)')
	loop.release()
	fp.finish()
	return epochs


def _project_pipelined(gv, reference_field, data_generator, test_data_generator, boundary_generator, boundary_lambda, batch_size, max_epoch, patience,
					   verbose, check_iter, history, rank, world, lattice_world, sample_seed, census, probe, use_graph):
	"""project() on the captured, pipelined iteration (timestep3d.ShardedProjector).  The projector of a (field, previous field)
	pair — buffers, optimiser state, the captured graph — is kept on the field object and reused by every later time step."""
	from . import timestep3d
	cur = reference_field.velocity_field
	box = data_generator.gsr_sampler[1]
	bbox = boundary_generator.gsr_sampler[1] if boundary_generator is not None else None
	lam = float(boundary_lambda) if boundary_generator is not None else 0.
	key = (id(cur), gv.N, int(batch_size), lam, box, bbox, float(reference_field.time_step), rank, world, int(sample_seed))
	cache = gv.__dict__.setdefault('_pipelines', {})
	fp = cache.get(key)
	if fp is None:
		cache.clear()	# one live projector per field: its buffers are sized by N and its graph holds this pair's pointers
		fp = timestep3d.ShardedProjector(gv, reference_field, lam, gv.N, int(batch_size), world=world, rank=rank, box=box, boundary_box=bbox, seed=int(sample_seed))
		cache[key] = fp
	else:
		fp.ref = reference_field
		fp.restart()
	fp.lattice_world = lattice_world
	names = ('loss_vor', 'loss_hel', 'loss_div')
	test = test_data_generator(gv)
	total = getattr(test_data_generator, 'total', None)
	fixed = bool(getattr(test_data_generator, 'fixed_points', False))
	if verbose:
		t = fp.evaluate_global(test, total, fixed=fixed).tolist()
		print(f'[projection] loss_vor: {t[0]}, loss_hel: {t[1]}, loss_div: {t[2]}')
	from .reseed import EarlyStop
	stop = EarlyStop(names, patience, check_iter)
	st_time = time.time()
	fp.begin(census)
	epochs, done = max_epoch, 0
	while done < max_epoch:
		n = min(check_iter, max_epoch - done)
		fp.run_iterations(n, census=census, use_graph=use_graph)
		done += n
		if n < check_iter:
			break
		t = fp.evaluate_global(test_data_generator(gv), total, probe=probe, census=census, fixed=fixed).tolist()	# the only host sync of the loop
		cur_l = dict(zip(names, t[:3]))
		if history is not None:
			history.setdefault('test', []).append(cur_l)
			history.setdefault('state', []).append(fp.stepper.scalars()[:16])
		if verbose:
			print(f'[projection] loss_vor: {t[0]}, loss_hel: {t[1]}, loss_div: {t[2]}, time: {time.time() - st_time}')
			st_time = time.time()
		if stop.update(cur_l):
			epochs = done
			if verbose:
				print('[projection] Total epoch:', epochs)
			break
	else:
		if verbose:
			print('[projection] Total epoch:', max_epoch, '(Reached maximum iteration number)')
	fp.finish()
	return epochs


def advance_frame(gaussian_velocity, new_gaussian_velocity, x_min, x_max, y_min, y_max, z_min, z_max, time_step, data_generator, test_data_generator,
				  boundary_generator=None, boundary_lambda=0., frame_id=None, max_epoch=20000, patience=500, verbose=1, fields=True, **project_kw):
	"""
	One pass of the reference's time loop (3D/advance.py:383-387, :389-390): clone (split over-stretched Gaussians), advect the
	Gaussians, project, swap; then the two output fields of the frame on the test lattice — |vorticity| and divergence, what the
	reference hands to write_vti.  Returns (gaussian_velocity, new_gaussian_velocity, epochs, (vorticity_norm, divergence) | None):
	the first is the field of the new frame.
	"""
	gv, new = gaussian_velocity, new_gaussian_velocity
	clone_velocity_field(new, gv, x_min, x_max, y_min, y_max, z_min, z_max, data_generator, test_data_generator, reinitialize=False, max_epoch=20000, verbose=verbose)
	advect_covector_field(new, gv, time_step, new.x_min, new.x_max, new.y_min, new.y_max, new.z_min, new.z_max)
	epochs = project(new, AdvectedCovectorField(gv, gv, time_step, x_min, x_max, y_min, y_max, z_min, z_max), x_min, x_max, y_min, y_max, z_min, z_max,
					 data_generator, test_data_generator, boundary_lambda=boundary_lambda, boundary_generator=boundary_generator, max_epoch=max_epoch,
					 patience=patience, verbose=verbose, frame_id=frame_id, **project_kw)
	gv, new = new, gv
	out = None
	if fields:
		lattice = test_data_generator(gv)
		grad = gv.gradient(lattice)	# ONE pass for both fields (the reference evaluates it once per written field, :374-377, :389-390: same values)
		vor = curl(grad).norm(dim=-1)
		div = grad.diagonal(dim1=-2, dim2=-1).sum(dim=-1)
		del grad
		census = project_kw.get('census')
		if census is not None:
			census.count(gv._engine, lattice, 1, lattice=True)
		out = (vor, div)
	return gv, new, epochs, out


def advance(gaussian_velocity, new_gaussian_velocity, x_min, x_max, y_min, y_max, z_min, z_max, time_step, last_time, boundary_generator=None,
			boundary_lambda=0., visualize_res=(128, 128, 128), out_dir=None, start_frame=0, max_epoch=20000, patience=500, verbose=1, on_frame=None,
			**project_kw):
	"""
	The time loop of 3D/advance.py:381-393: frames until `last_time`, each one advance_frame() with the stock generators; with
	`out_dir` the frame's vorticity_<k>.vti, divergence_<k>.vti and gaussian_velocity_<k>.pt are written as the reference does;
	`on_frame(k, field, vorticity_norm, divergence)` is called after every frame.  Returns the two field objects (current first).
	"""
	import os
	data_gen = BoxSampler(x_min, x_max, y_min, y_max, z_min, z_max)
	test_gen = LatticeGenerator(x_min, x_max, y_min, y_max, z_min, z_max, *visualize_res)
	gv, new = gaussian_velocity, new_gaussian_velocity
	seed0 = int(project_kw.pop('sample_seed', 42))
	seeds = {id(gv): seed0, id(new): seed0 + 1}	# the two fields alternate as the optimised one: each keeps its own sample stream
	t, cnt = 0., start_frame + 1
	while t < last_time:
		gv, new, _, (vor, div) = advance_frame(gv, new, x_min, x_max, y_min, y_max, z_min, z_max, time_step, data_gen, test_gen, boundary_generator=boundary_generator,
											   boundary_lambda=boundary_lambda, frame_id=cnt, max_epoch=max_epoch, patience=patience, verbose=verbose,
											   sample_seed=seeds[id(new)], **project_kw)
		if verbose:
			print(f'Wrote frame {cnt}')
		if out_dir is not None:
			gsr3d.write_vti_array(vor.reshape(visualize_res), x_min, x_max, y_min, y_max, z_min, z_max, os.path.join(out_dir, f'vorticity_{cnt}.vti'))
			gsr3d.write_vti_array(div.reshape(visualize_res), x_min, x_max, y_min, y_max, z_min, z_max, os.path.join(out_dir, f'divergence_{cnt}.vti'))
			gv.save(os.path.join(out_dir, f'gaussian_velocity_{cnt}.pt'))
		if on_frame is not None:
			on_frame(cnt, gv, vor, div)
		cnt += 1
		t += time_step
	return gv, new


def _project_unfused(gv, reference_field, data_generator, test_data_generator, boundary_generator, boundary_lambda,
					 batch_size, max_epoch, patience, verbose, check_iter, history):
	"""the reference's iteration structure on the CUDA kernels: explicit accumulators, host-side PCGrad, autograd regularisers, torch.optim"""
	device = _dev()
	W = PROJECT_WEIGHTS
	names = ('positions', 'scalings', 'rotations', 'values')

	def pcgrad_(g1, g2):
		if (g1 * g2).sum() < 0.:
			n1, n2 = g1 / g1.norm(), g2 / g2.norm()
			g1 -= (g1 * n2).sum() * n2
			g2 -= (g2 * n1).sum() * n1

	def get_losses(data, backward=True):
		ref_vor, ref_hel = reference_field.vorticity(data, need_hel=True)
		if backward:
			sets = {f'{tag}_{nm}_grad': torch.zeros_like(getattr(gv, nm)) for tag in ('vor', 'div') for nm in names}
			val, grad = gv.get_losses(data, ref_vor=ref_vor, weight_vor=W['vor'], ref_hel=ref_hel, weight_hel=W['hel'], weight_div=W['div'], **sets)
			for nm in names:
				g1, g2 = sets[f'vor_{nm}_grad'], sets[f'div_{nm}_grad']
				pcgrad_(g1, g2)
				getattr(gv, nm).grad += g1 + g2
		else:
			grad, val = gv.gradient(data, need_val=True)
		vor = curl(grad)
		loss_vor = (vor - ref_vor).abs().mean(dim=-1)
		loss_hel = ((val * vor).sum(dim=-1) - ref_hel).abs()
		loss_div = (grad[:, 0, 0] + grad[:, 1, 1] + grad[:, 2, 2]) ** 2
		loss_aniso, loss_vol = _regularisers(gv.scalings)
		loss_val_reg = gv.values.abs().mean()
		if backward:
			(W['aniso'] * loss_aniso + W['vol'] * loss_vol + W['val_reg'] * loss_val_reg).backward()
		boundary_constraint = torch.tensor(0., device=device)
		if boundary_lambda and boundary_generator:
			bdata, bnormal = boundary_generator(batch_size)
			if backward:
				bout, _ = gv.get_losses(bdata, normals=bnormal, weight_boundary=boundary_lambda)
			else:
				bout = gv(bdata)
			boundary_constraint = (bout * bnormal).sum(dim=1).abs().mean()
		loss_tot = (W['vor'] * loss_vor.mean() + W['div'] * loss_div.mean() + W['aniso'] * loss_aniso + W['vol'] * loss_vol
					+ W['val_reg'] * loss_val_reg + boundary_lambda * boundary_constraint)
		return loss_tot, loss_vor, loss_hel, loss_div

	keys = ('loss_vor', 'loss_hel', 'loss_div')
	best = {k: np.inf for k in keys}
	stale = {k: 0 for k in keys}
	epochs = max_epoch
	for epoch in range(max_epoch):
		data = data_generator(batch_size, gv)
		loss_tot, loss_vor, loss_hel, loss_div = get_losses(data)
		gv.step(loss_tot)
		if history is not None:
			history.setdefault('loss_tot', []).append(loss_tot.item())
		if epoch % check_iter == check_iter - 1:
			_, loss_vor, loss_hel, loss_div = get_losses(test_data_generator(gv), backward=False)
			cur = {'loss_vor': loss_vor.mean().item(), 'loss_hel': loss_hel.mean().item(), 'loss_div': loss_div.mean().item()}
			if verbose:
				print(f'[projection] {cur}')
			for k in keys:
				if cur[k] < best[k] * (1. - 1e-3):
					best[k], stale[k] = cur[k], 0
				else:
					stale[k] += check_iter
			if all(stale[k] >= patience for k in keys):
				epochs = epoch + 1
				break
	return epochs
