"""
2D time stepper — the roles of the reference's 2D/advance.py (AdvectedCovectorField :9-56, clone_velocity_field :58-158,
advect_covector_field :160-185, project :187-302) and of the fit loop of 2D/initialize.py:10-41, on the CUDA engine.

As in advance3d.py the optimisation loops exist in two forms with the same arithmetic:
  fused=True   device-resident iteration: forward, atomics-free backward, PCGrad + closed-form regularisers + Adam x4 +
               ReduceLROnPlateau + hash rebuild through gsr_step_rebuild — no host synchronisation inside an iteration;
  fused=False  the reference's own formulation on the drop-in class (get_grad_losses / get_losses, torch autograd
               regularisers, torch.optim.Adam), kept as the readable specification and as the parity partner in the tests.
The scenario (domains, scale factor, samplers) is passed explicitly (init_cond2d.Scene2D) instead of being read from
module globals keyed by the command line.
"""
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import gsr2d
from .engine import FusedStepper
from .graphloop import GraphedLoop


def _dev():
	return gsr2d.device


class AdvectedCovectorField:
	"""vorticity of `origin_covector_field` advected by `velocity_field` over `time_step` (2D/advance.py:9-56)"""

	def __init__(self, origin_covector_field, velocity_field, time_step, advection_scheme='rk4', domain=None):
		self.origin_covector_field, self.velocity_field = origin_covector_field, velocity_field
		self.time_step, self.advection_scheme = time_step, advection_scheme
		self.domain = domain	# (x_min, x_max, y_min, y_max) in GSR space: back-traced points outside it carry no vorticity (:52-53)

	def vorticity(self, x):
		if self.advection_scheme == 'rk4':
			return self.velocity_field.advected_vorticity(x, self.time_step, domain=self.domain)	# fused RK4 back-trace + curl
		if self.advection_scheme == 'rk1-backtrace':	# 2D/advance.py:34-44
			with torch.no_grad():
				xb = x - self.velocity_field(x) * self.time_step
				dv = self.velocity_field.gradient(xb.contiguous())
				vor = dv[:, 1, 0] - dv[:, 0, 1]
				if self.domain is not None:
					x0, x1, y0, y1 = self.domain
					vor = torch.where((xb[:, 0] < x0) | (xb[:, 0] > x1) | (xb[:, 1] < y0) | (xb[:, 1] > y1), torch.zeros_like(vor), vor)
			return vor
		raise NotImplementedError


def _regularisers(scalings, mask=None):
	"""anisotropy hinge at ratio 1.5 and volume uniformity (2D/advance.py:250-254)"""
	s = scalings if mask is None else scalings[mask]
	ratio = torch.exp(s.max(dim=-1).values - s.min(dim=-1).values) if s.shape[0] else torch.ones((1,), device=scalings.device)
	loss_aniso = (torch.where(ratio >= 1.5, ratio, torch.full_like(ratio, 1.5)) - 1.5).mean()
	return loss_aniso


def clone_velocity_field(res, velocity_field, data_generator, test_data_generator, batch_size=512, max_epoch=3000, patience=500, verbose=1, normals=None, seed=0):
	"""
	Copy velocity_field into res, split Gaussians with axis ratio >= 1.5 into two samples of their own distribution (once) and fit
	the new ones — and their neighbours — to the old field while everything else stays frozen (2D/advance.py:58-158).  The split
	runs on the device (reseed.py / csrc/split.cu); `normals` ((2, n_split, 2) standard-normal draws) and `seed` reproduce a split.
	"""
	from . import reseed
	with torch.no_grad():
		for nm in ('positions', 'scalings', 'rotations', 'values'):
			src, dst = getattr(velocity_field, nm).detach(), getattr(res, nm)
			if dst.shape == src.shape and dst.device == src.device:
				dst.copy_(src)	# same storage: a project() graph captured on these tensors stays valid for the next frame
			else:
				setattr(res, nm, src.clone())
		res.N = res.positions.shape[0]
		stop_gradient, n_split = reseed.split_all(res, 2, normals=[normals] if normals is not None else None, seed=seed, rounds=1)
	res.unfreeze()
	res.zero_grad()
	if n_split == 0:
		return res
	stop_gradient = torch.logical_and(stop_gradient, ~res.get_all_neighbors(res.positions[~stop_gradient].detach().contiguous()))
	if verbose:
		print(f'[clone] Add {n_split} particles.')
	sg = stop_gradient.int()

	def losses(data, backward):
		ref_grad, ref_val = velocity_field.gradient(data, need_val=True)
		if backward:
			val = res.get_losses(data, ref=ref_val, weight=1., stop_gradient=sg)
			grad = res.get_grad_losses(data, ref_grad=ref_grad, weight_grad=1., stop_gradient=sg)
		else:
			grad, val = res.gradient(data, need_val=True)
		loss, loss_grad = F.l1_loss(val, ref_val), F.l1_loss(grad, ref_grad)
		loss_aniso = _regularisers(res.scalings, ~stop_gradient)
		volumes = torch.where(stop_gradient, torch.exp(-res.scalings.detach().sum(dim=-1)), torch.exp(-res.scalings.sum(dim=-1)))
		loss_vol = ((volumes / volumes.mean() - 1) ** 2).mean()
		if backward:
			(loss_aniso + loss_vol).backward()
		return loss + loss_grad + loss_aniso + loss_vol, loss, loss_grad

	res.set_lr(positions_lr=1e-2, rotations_lr=5e-2, scalings_lr=5e-2, values_lr=5e-3)	# 2D/advance.py:120
	res.initialize_optimizers()
	reseed.refit(res, losses, data_generator, test_data_generator, ~stop_gradient, batch_size, max_epoch, patience, verbose)
	return res


def advect_covector_field(covector_field, velocity_field, dt, advection_scheme='rk4', extra_advector=None):
	"""move the Gaussians along the field and DROP those that leave the extended domain (2D/advance.py:160-185; note that the
	reference advects with covector_field's own field, :166)"""
	with torch.no_grad():
		if advection_scheme == 'rk1-backtrace':
			new_positions = covector_field.positions + dt * velocity_field(covector_field.positions.detach())
		elif advection_scheme == 'rk4':
			new_positions = covector_field.advection_rk4(covector_field.positions.detach(), dt)
		else:
			raise NotImplementedError
		valid = ((covector_field.x_min <= new_positions[:, 0]) & (new_positions[:, 0] <= covector_field.x_max)
				 & (covector_field.y_min <= new_positions[:, 1]) & (new_positions[:, 1] <= covector_field.y_max))
		if bool(valid.all()):	# nobody left the domain (the usual frame): move in place, the tensors — and graphs captured on them — stay
			covector_field.positions.detach().copy_(new_positions)
			params = None
		else:
			params = [new_positions[valid].clone(), covector_field.scalings.detach()[valid].clone(), covector_field.rotations.detach()[valid].clone(),
					  covector_field.values.detach()[valid].clone()]
	if params is not None:
		for p in params:
			p.requires_grad_()
		covector_field.positions, covector_field.scalings, covector_field.rotations, covector_field.values = params
	covector_field.N = covector_field.positions.shape[0]
	covector_field.zero_grad()
	if extra_advector:
		extra_advector(dt, advection_scheme)


PROJECT_WEIGHTS = dict(vor=1., div=1., aniso=10., vol=10., delta_pos=.5)	# 2D/advance.py:198
PROJECT_LRS = dict(positions=1e-4, scalings=1e-4, rotations=1e-4, values=1e-4)	# 2D/advance.py:261
KARMAN_LR_RATIO = 1.201956	# 2D/initialize.py:125, :163
INIT_PROJECT_WEIGHTS = dict(vor=1., div=10., aniso=10., vol=10., delta_pos=0.)	# 2D/initialize.py:55
INIT_PROJECT_LRS = dict(positions=1e-4, scalings=1e-5, rotations=1e-5 * KARMAN_LR_RATIO, values=1e-4)	# 2D/initialize.py:126


class FusedProjector2D:
	"""device-resident `project` phase (2D/advance.py:187-259 + GSR.step), one gsr_step_rebuild per iteration"""

	def __init__(self, gaussian_velocity, reference_field, boundary_lambda=0., patience=50, weights=None, lrs=None):
		gv = self.gv = gaussian_velocity
		self.ref = reference_field
		self.w = dict(PROJECT_WEIGHTS, **(weights or {}))
		self.lrs = dict(PROJECT_LRS, **(lrs or {}))
		self.boundary_lambda = float(boundary_lambda)
		e = gv._engine
		self.stepper = FusedStepper(e, [self.lrs[k] for k in ('positions', 'scalings', 'rotations', 'values')], patience,
									self.w['aniso'], self.w['vol'], w_dpos=self.w['delta_pos'], pcgrad=True,
									tau=gv.clamp_threshold, min_grid_scale=gv.min_grid_scale, ext_bounds=gv._ext())
		self.stepper.init(gv.scalings)
		e.build(gv.positions.detach(), params=[p.detach() for p in gv._params()])
		e._packed_key = None
		self.positions_org = gv.positions.detach().clone()
		cur = reference_field.velocity_field
		cur._engine.ensure_packed(cur._params())
		self._buf = {}
		self._streams = None

	def restart(self, reference_field):
		"""a new phase on the same tensors (the next frame): fresh optimiser state and hash; every buffer — and a graph captured on
		them — stays"""
		gv, e = self.gv, self.gv._engine
		self.ref = reference_field
		self.stepper.init(gv.scalings)
		e.build(gv.positions.detach(), params=[p.detach() for p in gv._params()])
		e._packed_key = None
		self.positions_org.copy_(gv.positions.detach())
		cur = reference_field.velocity_field
		cur._engine.ensure_packed(cur._params())

	def _tmp(self, name, shape):
		t = self._buf.get(name)
		if t is None or tuple(t.shape) != tuple(shape):
			t = torch.empty(shape, dtype=torch.float32, device=_dev())
			self._buf[name] = t
		return t

	def _ref_vorticity(self, data):
		"""self.ref.vorticity(data) into a persistent buffer (the result crosses streams: no allocator involvement)"""
		ref = self.ref
		if ref.advection_scheme != 'rk4':
			return ref.vorticity(data)
		cur = ref.velocity_field
		cur._engine.ensure_packed(cur._params())
		out = self._tmp('ref_vor', (data.shape[0],))
		cur._engine.advected_vorticity(data, -ref.time_step, out, None, domain=ref.domain)
		return out

	def iterate(self, data, boundary_1=None, boundary_2=None):
		"""
		One optimiser iteration; no host synchronisation.  Until the step the iteration is four independent chains — the RK4 pull-back
		reference (previous field), the forward pass of the training batch, and one {order, forward, adjoint + gather} chain per boundary
		batch — so they run on four streams (fork / join by events: a captured graph keeps the concurrency) and meet at the training
		gather and at the step.  Same kernels on the same inputs as the sequential order: same bits.
		"""
		gv, e = self.gv, self.gv._engine
		data = data.detach()
		Q = data.shape[0]
		lam = self.boundary_lambda
		main = torch.cuda.current_stream()
		if self._streams is None:
			self._streams = (torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream())
		# (forked only while being captured: eagerly the host is the bottleneck and the event calls would only slow it down)
		multi = torch.cuda.is_current_stream_capturing()
		s_fwd, s_b1, s_b2 = self._streams if multi else (main, main, main)
		fork = done_f = None
		if multi:
			fork = torch.cuda.Event()
			fork.record(main)
		srcs_b, extra, joins = [], [], []
		if lam > 0. and boundary_1 is not None:	# prescribed velocity on an obstacle (2D/advance.py:217-220)
			bdata, bvalue = [t.detach() for t in boundary_1]
			if multi:
				s_b1.wait_event(fork)
			with torch.cuda.stream(s_b1):
				bins1 = e.bin_samples(bdata, True, tag='b1')
				val1 = self._tmp('val1', (bdata.shape[0], 2))
				e.forward(bdata, val1, None, accumulate=False, perm=bins1)
				acc1, _ = e.backward_gather(bdata, bins1.perm, bins1.scs, val1, None, (lam, 0., 0., 0., 0., 0.), {'ref_val': bvalue}, None, tag='acc_b1', want_losses=True)
				lp1, nb1 = e.last_loss_partials
				if multi:
					ev = torch.cuda.Event()
					ev.record(s_b1)
					joins.append(ev)
			srcs_b.append((lp1, nb1, [0., 0., 0., 0., lam / bdata.shape[0], 0., 0., 0.]))
			extra.append(acc1)
		if lam > 0. and boundary_2 is not None:	# prescribed normal velocity (2D/advance.py:231-235)
			bdata2, bnormal, bref = [t.detach() for t in boundary_2]
			if multi:
				s_b2.wait_event(fork)
			with torch.cuda.stream(s_b2):
				bins2 = e.bin_samples(bdata2, True, tag='b2')
				val2 = self._tmp('val2', (bdata2.shape[0], 2))
				e.forward(bdata2, val2, None, accumulate=False, perm=bins2)
				acc2, _ = e.backward_gather(bdata2, bins2.perm, bins2.scs, val2, None, (0., lam, 0., 0., 0., 0.), {'normals': bnormal, 'normal_ref': bref}, None,
											tag='acc_b2', want_losses=True)
				lp2, nb2 = e.last_loss_partials
				if multi:
					ev = torch.cuda.Event()
					ev.record(s_b2)
					joins.append(ev)
			srcs_b.append((lp2, nb2, [0., 0., 0., lam / bdata2.shape[0], 0., 0., 0., 0.]))
			extra.append(acc2)
		bins = e.bin_samples(data, True)
		grad = self._tmp('grad', (Q, 2, 2))
		if multi:
			binned = torch.cuda.Event()
			binned.record(main)
			s_fwd.wait_event(binned)
		with torch.cuda.stream(s_fwd):
			e.forward(data, None, grad, accumulate=False, perm=bins)
			if multi:
				done_f = torch.cuda.Event()
				done_f.record(s_fwd)
		ref_vor = self._ref_vorticity(data)	# beside the forward pass
		if multi:
			main.wait_event(done_f)
		acc, mask = e.backward_gather(data, bins.perm, bins.scs, None, grad, (0., 0., 0., self.w['vor'], 0., self.w['div']), {'ref_vor': ref_vor}, None, want_losses=True)
		lp, nblk = e.last_loss_partials
		srcs = [(lp, nblk, [self.w['vor'] / Q, 0., self.w['div'] / Q, 0., 0., 0., 0., 0.])] + srcs_b
		for ev in joins:
			main.wait_event(ev)
		self.stepper.step([p.detach() for p in gv._params()], acc, mask, extra=extra, loss_srcs=srcs, positions_org=self.positions_org, rebuild=True)

	def evaluate(self, data):
		"""(mean |omega - omega_ref|, mean (div u)^2) on `data`, device tensor of 2"""
		gv, e = self.gv, self.gv._engine
		data = data.detach()
		Q = data.shape[0]
		ref_vor = self.ref.vorticity(data)
		grad = self._tmp('t_grad', (Q, 2, 2))
		e.forward(data, None, grad, accumulate=False, perm=e.bin_samples(data, False, tag='t'))
		s = e.sample_losses(None, grad, {'ref_vor': ref_vor}, Q) / Q
		return torch.stack([s[0], s[2]])

	def finish(self):
		gv = self.gv
		gv.grid_scale = self.stepper.detach()
		gv._engine._packed_key = None
		for p in gv._params():
			p.add_(0.)	# the tensors were modified through raw pointers: bump the autograd version counters
		gv._engine.set_grid(gv._ext(), gv.grid_size, gv.grid_scale, gv.clamp_threshold)
		gv._engine.build(gv.positions.detach())


def _project_unfused(gv, reference_field, data_generator, test_data_generator, b1, b2, lam, batch_size, max_epoch, patience, verbose, check_iter, weights=None):
	"""the reference's formulation, statement by statement (2D/advance.py:187-302)"""
	device = _dev()
	positions_org = gv.positions.detach().clone()
	w = dict(PROJECT_WEIGHTS, **(weights or {}))

	def gradient_project(g1, g2):
		if (g1 * g2).sum() < 0.:
			n1, n2 = g1 / (g1 ** 2).sum() ** .5, g2 / (g2 ** 2).sum() ** .5
			g1 -= (g1 * n2).sum() * n2
			g2 -= (g2 * n1).sum() * n1

	def get_losses(data, backward=True):
		ref_vor = reference_field.vorticity(data)
		bc = torch.tensor(0., device=device)
		if backward:
			sets = {f'{t}_{n}_grad': torch.zeros_like(getattr(gv, n)) for t in ('vor', 'div') for n in ('positions', 'scalings', 'rotations', 'values')}
			grad = gv.get_grad_losses(data, ref_vor=ref_vor, weight_vor=w['vor'], weight_div=w['div'], **sets)
			if lam > 0. and b1:
				bd, bv = b1(batch_size)
				bc = bc + F.l1_loss(gv.get_losses(bd, ref=bv, weight=lam), bv)
			for n in ('positions', 'scalings', 'rotations', 'values'):
				gradient_project(sets[f'vor_{n}_grad'], sets[f'div_{n}_grad'])
				getattr(gv, n).grad += sets[f'vor_{n}_grad'] + sets[f'div_{n}_grad']
			if lam > 0. and b2:
				bd, bn, bref = b2(batch_size)
				out = gv.get_losses(bd, normals=bn, normal_ref=bref, weight_boundary=lam)
				bc = bc + F.l1_loss((out * bn).sum(dim=1), bref)
		else:
			grad = gv.gradient(data)
		loss_vor = torch.abs(grad[:, 1, 0] - grad[:, 0, 1] - ref_vor)
		loss_div = (grad[:, 0, 0] + grad[:, 1, 1]) ** 2
		loss_aniso = _regularisers(gv.scalings)
		volumes = torch.exp(-gv.scalings.sum(dim=-1))
		loss_vol = ((volumes / volumes.mean() - 1) ** 2).mean()
		loss_dpos = F.mse_loss(gv.positions, positions_org)
		if backward:
			(w['aniso'] * loss_aniso + w['vol'] * loss_vol + w['delta_pos'] * loss_dpos).backward()
		tot = w['vor'] * loss_vor.mean() + w['div'] * loss_div.mean() + w['aniso'] * loss_aniso + w['vol'] * loss_vol + w['delta_pos'] * loss_dpos + lam * bc
		return tot, loss_vor.mean(), loss_div.mean()

	best, stale = [np.inf, np.inf], [0, 0]
	for epoch in range(max_epoch):
		tot, _, _ = get_losses(data_generator(batch_size, gv))
		gv.step(tot)
		if epoch % check_iter == check_iter - 1:
			with torch.no_grad():
				_, lv, ld = get_losses(test_data_generator(gv), backward=False)
			for k, (v, thr) in enumerate(((lv.item(), 1e-3), (ld.item(), 1e-2))):
				if v < best[k] * (1. - thr):
					best[k], stale[k] = v, 0
				else:
					stale[k] += check_iter
			if verbose:
				print(f'[projection] loss_vor: {lv.item()}, loss_div: {ld.item()}')
			if stale[0] >= patience and stale[1] >= patience:
				return epoch + 1
	return max_epoch


def project(gaussian_velocity, reference_field, data_generator, test_data_generator, boundary_generator_1=None, boundary_generator_2=None,
			boundary_lambda=0., batch_size=512, max_epoch=3000, patience=500, verbose=1, fused=True, check_iter=100, weights=None, lrs=None, use_graph=None,
			cache=True):
	"""
	One time step's projection by first-order optimisation (2D/advance.py:187-302): match the advected vorticity, drive
	the divergence to zero, keep the Gaussians well shaped and close to their advected positions.  Early stop: every 100
	iterations, vorticity must improve by 0.1 % or divergence by 1 %, `patience` iterations without either ends the phase.
	Returns the number of iterations run.  `weights` / `lrs` override PROJECT_WEIGHTS / PROJECT_LRS: the reference keeps a second
	copy of this function with other constants for the Karman initial state (2D/initialize.py:44-160, INIT_PROJECT_* below).
	use_graph: replay the iterations as a CUDA graph (graphloop.py).  That is only valid when the generators are pure device
	functions of torch's CUDA random stream — a generator that walks a Python list would be replayed with its first batches — so
	the default (None) turns it on only when every generator carries `graph_safe = True` (Scene2D's samplers do).
	cache: keep the projector and its captured graph on the field object and reuse them in the next frame when NOTHING the graph
	holds has changed — the same tensors (clone and advect update in place when no Gaussian was split or dropped), the same
	reference field object, time step, domain, weights, rates and generator objects — instead of warming up and capturing again.
	"""
	gv = gaussian_velocity
	lr = dict(PROJECT_LRS, **(lrs or {}))
	gv.set_lr(positions_lr=lr['positions'], scalings_lr=lr['scalings'], rotations_lr=lr['rotations'], values_lr=lr['values'])
	gv.initialize_optimizers(patience=50)
	for s in gv.schedulers:
		s.factor = .9
	if not fused:
		return _project_unfused(gv, reference_field, data_generator, test_data_generator, boundary_generator_1, boundary_generator_2, boundary_lambda,
								batch_size, max_epoch, patience, verbose, check_iter, weights)
	use_b1 = boundary_lambda > 0. and boundary_generator_1
	use_b2 = boundary_lambda > 0. and boundary_generator_2
	if use_graph is None:
		use_graph = all(getattr(g, 'graph_safe', False) for g in (data_generator, boundary_generator_1 if use_b1 else data_generator, boundary_generator_2 if use_b2 else data_generator))
	cur = reference_field.velocity_field
	key = None
	if cache and use_graph:
		key = (id(cur), gv.N, cur.N, float(reference_field.time_step), tuple(float(v) for v in (reference_field.domain or ())), reference_field.advection_scheme,
			   float(boundary_lambda), batch_size, check_iter, tuple(sorted((weights or {}).items())), tuple(sorted(lr.items())),
			   id(data_generator), id(boundary_generator_1) if use_b1 else 0, id(boundary_generator_2) if use_b2 else 0,
			   tuple(p.data_ptr() for p in gv._params()), tuple(p.data_ptr() for p in cur._params()))
		store = gv.__dict__.setdefault('_pipelines2d', {})
		hit = store.get(key)
		if hit is not None:
			fp, loop = hit
			fp.restart(reference_field)
		else:
			for old_fp, old_loop in store.values():	# tensors or settings changed: the old graphs can never be replayed again
				old_loop.release()
			store.clear()
	if key is None or hit is None:
		fp = FusedProjector2D(gv, reference_field, boundary_lambda, patience=50, weights=weights, lrs=lrs)

	def batches():	# sync-free: the draws come from torch's CUDA generator
		data = data_generator(batch_size, gv)
		b1 = boundary_generator_1(batch_size) if use_b1 else None
		b2 = boundary_generator_2(batch_size) if use_b2 else None
		return data, b1, b2

	def iteration(inputs):	# sync-free: every scalar that changes between iterations lives in the stepper's device state
		fp.iterate(*inputs)
	if key is None or hit is None:
		loop = GraphedLoop(iteration, unit=next(u for u in (10, 5, 2, 1) if check_iter % u == 0), enabled=use_graph, prepare=batches)
		if key is not None:
			store[key] = (fp, loop)
	best, stale = [np.inf, np.inf], [0, 0]
	epochs = max_epoch
	st_time = time.time()
	done = 0
	while done < max_epoch:
		k = min(check_iter, max_epoch - done)
		loop.run(k)
		done += k
		if k < check_iter:
			break
		lv, ld = fp.evaluate(test_data_generator(gv)).tolist()	# the only host synchronisation: once per check_iter iterations
		for j, (v, thr) in enumerate(((lv, 1e-3), (ld, 1e-2))):
			if v < best[j] * (1. - thr):
				best[j], stale[j] = v, 0
			else:
				stale[j] += check_iter
		if verbose:
			print(f'[projection] loss_vor: {lv}, loss_div: {ld}, time: {time.time() - st_time}')
			st_time = time.time()
		if stale[0] >= patience and stale[1] >= patience:
			epochs = done
			break
	if key is None:
		loop.release()
	fp.finish()
	return epochs


def fit_velocity_with_gradient(gaussian_velocity, reference_field, reference_gradient, data_generator, batch_size=512, max_epoch=3000, verbose=1, fused=True, use_graph=None):
	"""
	Initial fit of the representation to an analytic field: value L1 + gradient L1 + anisotropy + volume regularisers
	(2D/initialize.py:10-41).  fused=True runs the iteration through gsr_step_rebuild (no PCGrad: a single gradient set).
	"""
	gv = gaussian_velocity
	gv.initialize_optimizers()
	if not fused:
		for epoch in range(max_epoch):
			data = data_generator(batch_size)
			ref_val, ref_grad = reference_field(data), reference_gradient(data)
			val = gv.get_losses(data, ref=ref_val, weight=1.)
			grad = gv.get_grad_losses(data, ref_grad=ref_grad, weight_grad=1.)
			volumes = torch.exp(-gv.scalings.sum(dim=-1))
			loss_aniso, loss_vol = _regularisers(gv.scalings), ((volumes / volumes.mean() - 1) ** 2).mean()
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
		return data, reference_field(data).contiguous(), reference_gradient(data).contiguous()

	def iteration(inputs):
		data, ref_val, ref_grad = inputs
		Q = data.shape[0]
		bins = e.bin_samples(data, True)
		val, grad = torch.empty((Q, 2), device=_dev()), torch.empty((Q, 2, 2), device=_dev())
		e.forward(data, val, grad, accumulate=False, perm=bins)
		# value and gradient losses live in the same (direct) accumulator set: one gather serves both (the reference runs two kernels)
		acc, mask = e.backward_gather(data, bins.perm, bins.scs, val, grad, (1., 0., 1., 0., 0., 0.), {'ref_val': ref_val, 'ref_grad': ref_grad}, None, want_losses=True)
		lp, nblk = e.last_loss_partials
		stepper.step([p.detach() for p in gv._params()], acc, mask, loss_srcs=[(lp, nblk, [0., 0., 0., 0., 1. / Q, 1. / Q, 0., 0.])], rebuild=True)
	if use_graph is None:
		use_graph = all(getattr(g, 'graph_safe', False) for g in (data_generator, reference_field, reference_gradient))
	loop = GraphedLoop(iteration, unit=10, enabled=use_graph, prepare=batches)
	done = 0
	while done < max_epoch:
		k = min(100, max_epoch - done)
		loop.run(k)
		done += k
		if verbose and k == 100:
			sc = stepper.scalars()
			print(f'loss_tot: {sc[9]}, loss_aniso: {sc[10]}, loss_vol: {sc[11]}, time: {time.time() - st_time}')
			st_time = time.time()
	loop.release()
	gv.grid_scale = stepper.detach()
	e._packed_key = None
	for p in gv._params():
		p.add_(0.)
	e.set_grid(gv._ext(), gv.grid_size, gv.grid_scale, gv.clamp_threshold)
	e.build(gv.positions.detach())


def init_karman_velocity(gaussian_velocity, scene, reference_field, reference_gradient, data_generator, batch_size=512, max_epoch=3000, verbose=1, fused=True,
						 project_epochs=10000, use_graph=None):
	"""
	The Karman initial state (2D/initialize.py:162-185): fit the uniform inflow with ten times smaller scaling / rotation rates than
	the other scenes, then project the fit onto the divergence-free fields that respect the obstacle and the channel walls — a
	`project` against the field's own vorticity (dt = 0) with the initial-state constants (divergence weight 10, no position
	anchor, lrs 1e-4 / 1e-5 / 1.2e-5 / 1e-4), boundary weight 10, samples and test lattice on the advance domain, and an early stop
	that cannot trigger before `project_epochs` (patience = max_epoch = 10000 in the reference).
	"""
	gv = gaussian_velocity
	gv.set_lr(positions_lr=1.6e-4 * 10., scalings_lr=5e-3, rotations_lr=5e-3 * KARMAN_LR_RATIO, values_lr=5e-4 * 10.)
	fit_velocity_with_gradient(gv, reference_field, reference_gradient, data_generator, batch_size, max_epoch, verbose, fused=fused, use_graph=use_graph)
	x_min, x_max, y_min, y_max = scene.scaled(scene.initialize_domain)
	x_N, y_N = scene.particle_count
	tmp = gsr2d.GaussianSplattingFast(x_min, x_max, y_min, y_max, gsr2d.get_grid_points(x_min, x_max, y_min, y_max, x_N, y_N).cpu().numpy(), dim=2)
	with torch.no_grad():
		for nm in ('positions', 'scalings', 'rotations', 'values'):
			setattr(tmp, nm, getattr(gv, nm).detach().clone())
		tmp.N = tmp.positions.shape[0]
	tmp.unfreeze()
	tmp.zero_grad()
	b1, b2 = scene.boundary_samplers
	ref = AdvectedCovectorField(tmp, tmp, 0., domain=scene.scaled(scene.advance_domain))
	gen = lambda n, gs, restrict=None: scene.data_generator(gs)
	gen.graph_safe = True
	return project(gv, ref, gen, lambda gs: scene.test_generator(), boundary_generator_1=b1, boundary_generator_2=b2,
				   boundary_lambda=10., patience=project_epochs, max_epoch=project_epochs, verbose=verbose, fused=fused, weights=INIT_PROJECT_WEIGHTS, lrs=INIT_PROJECT_LRS, use_graph=use_graph)


def simulation_initialize(scene, max_epoch=10000, verbose=1, fused=True, project_epochs=10000, use_graph=None):
	"""SimulationInitialize of 2D/initialize.py:187-238 without the plots: lattice of Gaussians -> fit (karman: -> projection onto
	the boundary conditions) -> the frame-0 field"""
	x_min, x_max, y_min, y_max = scene.scaled(scene.initialize_domain)
	x_N, y_N = scene.particle_count
	pts = gsr2d.get_grid_points(x_min, x_max, y_min, y_max, x_N, y_N).cpu().numpy()
	gv = gsr2d.GaussianSplattingFast(x_min, x_max, y_min, y_max, pts, dim=2)
	gen = lambda n: scene.data_generator(gv, domain=scene.initialize_domain)
	gen.graph_safe = True
	if scene.name == 'karman':
		init_karman_velocity(gv, scene, scene.target_velocity, scene.target_gradient, gen, max_epoch=max_epoch, verbose=verbose, fused=fused, project_epochs=project_epochs, use_graph=use_graph)
		return gv
	gv.set_lr(positions_lr=1.6e-3, scalings_lr=5e-2, rotations_lr=5e-2, values_lr=5e-3)
	fit_velocity_with_gradient(gv, scene.target_velocity, scene.target_gradient, gen, max_epoch=max_epoch, verbose=verbose, fused=fused, use_graph=use_graph)
	return gv


def advance(scene, gaussian_velocity, new_gaussian_velocity, dt, max_epoch=20000, boundary_lambda=1., verbose=1, fused=True):
	"""one frame of the `while t < last_time` loop of 2D/advance.py:354-365; returns (current, spare) after the swap"""
	gens = scene.__dict__.get('_advance_generators')	# the same generator OBJECTS every frame: project() keeps its captured graph across frames
	if gens is None:
		gen = lambda n, gs, restrict=None: scene.data_generator(gs)
		gen.graph_safe = True
		gens = scene._advance_generators = (gen, lambda gs: scene.test_generator()) + tuple(scene.boundary_samplers)
	gen, test, b1, b2 = gens
	clone_velocity_field(new_gaussian_velocity, gaussian_velocity, gen, test, max_epoch=max_epoch, verbose=verbose)
	advect_covector_field(new_gaussian_velocity, gaussian_velocity, dt, extra_advector=scene.extra_advector)	# karman: the inlet moves with the flow
	ref = AdvectedCovectorField(gaussian_velocity, gaussian_velocity, dt, domain=scene.scaled(scene.advance_domain))
	project(new_gaussian_velocity, ref, gen, test, boundary_generator_1=b1, boundary_generator_2=b2, boundary_lambda=boundary_lambda, max_epoch=max_epoch,
			verbose=verbose, fused=fused)
	return new_gaussian_velocity, gaussian_velocity
