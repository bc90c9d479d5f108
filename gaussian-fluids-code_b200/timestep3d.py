"""
The pipelined execution of `advance3d.project` and the fixed-work 3D time step used for measurement.

ShardedProjector is what `advance3d.project()` runs on when it is given the stock generators: the samples are drawn on the
device (counter-based Philox kernels keyed by the optimiser state's running sample clock), ten iterations form one captured CUDA
graph, and everything that does not depend on the updated Gaussians is moved off the iteration's critical cycle — the next
iteration's samples and their hashes are prepared right after the step (beside the Gaussian hash), the RK4 pull-back reference of
those samples runs on its own stream until the next adjoint kernel needs it.

LeapfrogTimestep is one frame of `3D/advance.py` through that API (advance3d.advance_frame) with the iteration count pinned to the
reference's minimum, because the reference's own count is data dependent:

    clone (copy, no split)  ->  advect (RK4 positions only, Q = N)
    ->  `iters` (600) project iterations, each { RK4 pull-back reference (5 evaluations), forward + backward with the
        vorticity / helicity / divergence losses on Q = N samples, boundary forward + backward on 8192 samples,
        PCGrad + regularisers + Adam x4 + scheduler, hash rebuild }
    ->  a test pass { RK4 pull-back + forward on the test_res^3 lattice } every 100 iterations
    ->  swap, two forward passes on the lattice (the |vorticity| and divergence fields the reference writes as VTI).

Multi-GPU (one process per GPU), two modes:
  * scaling='weak': sample points are sharded — every rank draws its own N training and 8192 boundary samples (global
    Q = world * N, the loss normalisers use the global counts) and its own test_res^3 share of a lattice that is world times finer
    along z (fixed work per rank); Gaussian parameters, hash and optimiser state are replicated; ONE exchange per iteration sums
    the compact gradient accumulators and the loss partial sums (a single kernel over NVLink peer memory, csrc/xrank.cu; NCCL
    all-reduce as the fallback), after which every rank runs the identical fused step, so the replicas stay bit-identical without
    a broadcast.
  * scaling='strong': the reference's frame itself, made faster — the fixed test_res^3 lattice (80 % of the frame's sample
    evaluations at the reference's size) is split over the ranks by z planes, no collective on the data path; the 600
    training iterations on N = 1000 Gaussians are latency bound and run replicated (sharding them would only add an exchange to
    every 40 us iteration).
"""
import ctypes as C
import os

import torch

from . import _lib, advance3d, engine, gsr3d
from .init_cond3d import sample_on_box
from .synth import make_fast3d, synthetic_field


def shard_lattice(test_res, rank, world, device, scaling='weak'):
	"""
	This rank's share of the test / output lattice (world = 1: the reference's test_res^3 lattice, 3D/GSR.py:719-725).
	weak:   fixed work per rank — the job evaluates a test_res x test_res x (world * test_res) lattice, rank r owns its z planes
	        r, r + world, ...
	strong: fixed job — the test_res^3 lattice itself, rank r owns its z planes r, r + world, ...
	Returns (points, number of points of the whole job's lattice).
	"""
	ax = torch.linspace(0., 1., test_res, device=device)
	nz = test_res * world if scaling == 'weak' else test_res
	az = torch.linspace(0., 1., nz, device=device)[rank::world]
	return torch.stack(torch.meshgrid(ax, ax, az, indexing='ij'), dim=-1).reshape(-1, 3).contiguous(), test_res * test_res * nz


def flat_layout(N, nblk, nblkb, AF=12):
	"""offsets of the per-iteration all-reduce buffer: [3 accumulator sets (N, AF) | loss partials | boundary loss partials]"""
	a = 3 * N * AF
	return {'acc': (0, a), 'lp': (a, a + nblk * 8), 'lpb': (a + nblk * 8, a + (nblk + nblkb) * 8), 'total': a + (nblk + nblkb) * 8}


class Census:
	"""device-side counter of candidate visits (the benchmark's unit of work); the visits made on the test / output lattice are
	also kept apart (in the strong-scaling mode only those are shared between the ranks, the training visits are replicated)"""

	def __init__(self, device):
		self.c = torch.zeros(2, dtype=torch.int64, device=device)	# [candidate visits C, accepted pairs P], everything
		self.lat = torch.zeros(2, dtype=torch.int64, device=device)	# the lattice passes' share of it
		self.skipped = torch.zeros(2, dtype=torch.int64, device=device)	# NOT executed (and not in c): the reference's repeated pull-backs of the fixed lattice

	def count(self, engine, x, evals, lattice=False, executed=True):
		"""executed=False: visits the reference's algorithm makes at this point and this engine does not (the reused pull-back reference
		of the test lattice): kept apart in `skipped`, never part of the executed totals"""
		if not executed:
			engine.count_pairs(x, self.skipped, evals, True)
			return
		before = self.c.clone() if lattice else None
		engine.count_pairs(x, self.c, evals, True)
		if lattice:
			self.lat += self.c - before

	def skipped_value(self):
		c = self.skipped.tolist()
		return int(c[0]), int(c[1])

	def value(self):
		c = self.c.tolist()
		return int(c[0]), int(c[1])

	def lattice_value(self):
		c = self.lat.tolist()
		return int(c[0]), int(c[1])


class PeerExchange:
	"""
	The per-iteration exchange of the sharded optimisation as ONE kernel over NVLink peer memory (csrc/xrank.cu): every rank's
	accumulator buffer lives in symmetric memory (torch.distributed._symmetric_memory: CUDA VMM allocations mapped into every
	rank of the node), in two parities, and gsr_xrank_sum adds the peers' buffers in rank order after a signal-pad handshake.
	"""
	SIGNAL_WORD0 = 256	# first uint32 of the signal pad used here (torch's own barriers use the low words)

	def __init__(self, n_floats, device):
		import torch.distributed as dist
		import torch.distributed._symmetric_memory as symm_mem
		self.n = (int(n_floats) + 3) // 4 * 4
		self.rank, self.world = dist.get_rank(), dist.get_world_size()
		self.bufs, self.hdls = [], []
		for _ in range(2):
			b = symm_mem.empty(self.n, dtype=torch.float32, device=device)
			b.zero_()
			self.hdls.append(symm_mem.rendezvous(b, dist.group.WORLD))
			self.bufs.append(b)
		h = self.hdls[0]
		if h.signal_pad_size < 4 * (self.SIGNAL_WORD0 + self.world):
			raise _lib.GsrError('signal pad too small')
		self.out = torch.zeros(self.n, dtype=torch.float32, device=device)
		self.err = torch.zeros(1, dtype=torch.int32, device=device)
		self.base = torch.zeros(1, dtype=torch.int32, device=device)
		self._base_host = 0
		self._ptrs = [(C.c_void_p * self.world)(*[int(p) for p in hd.buffer_ptrs]) for hd in self.hdls]
		self._sigs = (C.c_void_p * self.world)(*[int(p) + 4 * self.SIGNAL_WORD0 for p in h.signal_pad_ptrs])
		torch.cuda.synchronize()
		dist.barrier()	# every rank's buffers are zeroed and mapped before anyone signals

	def new_phase(self):
		"""epochs are base + iteration + 1: a new optimisation phase (iteration counter back to 0) moves the base past the old ones;
		a timeout flag of the previous phase is reported here at the latest"""
		self.check()
		self._base_host += 100000
		self.base.fill_(self._base_host)

	def sum(self, parity, iteration_dev):
		lib = _lib.lib()
		_lib.check(lib.gsr_xrank_sum(self._ptrs[parity], self._sigs, C.c_int(self.rank), C.c_int(self.world), C.c_int64(self.n), _lib.ptr(iteration_dev),
									 _lib.ptr(self.base, torch.int32), _lib.ptr(self.out), _lib.ptr(self.err, torch.int32), _lib.stream()), 'gsr_xrank_sum')

	def check(self):
		if int(self.err.item()):
			raise _lib.GsrError('gsr_xrank_sum: a peer did not publish its buffer in time')


class ShardedProjector(advance3d.FusedProjector):
	"""FusedProjector whose accumulators and loss partials live in one flat buffer that is all-reduced once per iteration"""

	def __init__(self, gv, reference_field, boundary_lambda, Q, Qb, world=1, rank=0, patience=50, box=(0., 1.) * 3, boundary_box=None, seed=42):
		# samples of iteration k + 1 drawn and binned while iteration k still runs (see iterate()); GSR_SAMPLES_AHEAD=0: after step k
		self.ahead = os.environ.get('GSR_SAMPLES_AHEAD', '1') != '0'
		super().__init__(gv, reference_field, boundary_lambda, patience=patience)
		e = gv._engine
		self.world, self.rank = world, rank
		self.lattice_world = world	# processes sharing the test lattice (LeapfrogTimestep: also > 1 in the strong mode, where world == 1 here)
		self.Q, self.Qb = Q, Qb
		self.box, self.boundary_box, self.seed = tuple(box), tuple(boundary_box or box), seed
		dev = gsr3d.device
		# persistent sample buffers: the device samplers write them, the captured graph reads them
		self._x = torch.empty((Q, 3), dtype=torch.float32, device=dev)
		self._xb, self._nb = torch.empty((Qb, 3), dtype=torch.float32, device=dev), torch.empty((Qb, 3), dtype=torch.float32, device=dev)
		self._clock1 = torch.zeros(1, dtype=torch.float32, device=dev)	# sample clock + 1: the counter of a batch drawn one iteration early
		self.graph, self.unit, self.per_iter, self.graph_launches = None, 0, 0, 0
		N = gv.N
		nblk, nblkb = e.lib.gsr_loss_blocks(Q), e.lib.gsr_loss_blocks(Qb)
		lay = flat_layout(N, nblk, nblkb)
		# exchange: 'p2p' = one kernel over NVLink peer memory (default when sharded), 'nccl' = torch.distributed.all_reduce
		self.exchange = os.environ.get('GSR_EXCHANGE', 'p2p') if world > 1 else 'none'
		self.peer = None
		if self.exchange == 'p2p':
			try:
				self.peer = PeerExchange(lay['total'], gsr3d.device)
			except Exception as ex:	# no symmetric memory on this system: say so and use the library collective
				print(f'[gsr] peer-memory exchange unavailable ({type(ex).__name__}: {ex}); using NCCL all_reduce', flush=True)
				self.exchange = 'nccl'
		flats = self.peer.bufs if self.peer else [torch.zeros(lay['total'], dtype=torch.float32, device=gsr3d.device)]
		view = lambda f: (f[lay['acc'][0]:lay['acc'][1]].view(3, N, 12), f[lay['lp'][0]:lay['lp'][1]].view(nblk, 8), f[lay['lpb'][0]:lay['lpb'][1]].view(nblkb, 8))
		self.flats = flats
		self.views = [view(f) for f in flats]	# per parity: (acc, lp, lpb) the gather kernels write
		self.reduced = view(self.peer.out) if self.peer else self.views[0]	# what the fused step reads
		self.flat = flats[0]
		self.acc, self.lp, self.lpb = self.views[0]	# set 0: boundary (direct), sets 1, 2: vorticity / divergence
		self.nblk, self.nblkb = nblk, nblkb
		self._streams = None
		self._samplers = self._prep = self._ident = None
		if self.peer:
			self.peer.new_phase()
		self.set_samplers(self.draw_samples, self.draw_boundary_binned if self.boundary_lambda else None)

	# ---- device samplers (one kernel each; the iteration number comes from the state's running sample clock) ----------------------
	def draw_samples(self, clock=None):
		"""this rank's training samples of the coming iteration: uniform in the box, Q = N (3D/advance.py:339-340)"""
		return self.gv._engine.sample_box(self.box, self._x, self.seed, 2 * self.rank, clock if clock is not None else self.stepper.clock)

	def draw_boundary(self, clock=None, gs_dev=None):
		"""sample_on_box(Qb) (3D/init_cond.py:227-249) (unordered: gs_dev is for the samplers that also order the batch)"""
		return self.gv._engine.sample_box_surface(self.boundary_box, self._xb, self._nb, self.seed, 2 * self.rank + 1, clock if clock is not None else self.stepper.clock)

	def draw_boundary_binned(self, clock=None, gs_dev=None):
		"""draw_boundary + the engine's ordering of the batch, one launch: ((points, normals), Bins)"""
		bins = self.gv._engine.sample_box_surface_binned(self.boundary_box, self._xb, self._nb, self.seed, 2 * self.rank + 1,
														 clock if clock is not None else self.stepper.clock, tag='pb', gs_dev=gs_dev)
		return (self._xb, self._nb), bins

	def restart(self, keep_clock=True):
		"""begin a new project phase on the same buffers: fresh optimiser state, fresh hash; the sample clock runs on, so the new
		phase draws new samples (keep_clock=False: replay the first phase's samples)"""
		if self.peer:
			self.peer.new_phase()
		if getattr(self, '_drop_clock', False):	# LeapfrogTimestep.reset(): replay from the first sample
			keep_clock, self._drop_clock = False, False
		self.stepper.init(self.gv.scalings, keep_clock=keep_clock)
		self._it = 0
		self._test_ref_key = None	# the previous field has changed: the first test pass of the phase evaluates its pull-back reference anew
		self._rebuild()
		cur = self.ref.velocity_field
		cur._engine._packed_key = None
		cur._engine.ensure_packed(cur._params())

	# ---- sample pipeline ------------------------------------------------------------------------------------------------
	# Generating and binning the samples of iteration k+1 needs only the iteration counter and the grid_scale that step k
	# leaves in the device state — not the updated Gaussians.  So it is forked right after the step kernels and runs beside
	# the Gaussian hash + pack instead of in front of the next forward pass (the boundary batch's hash was 17 us of a 78 us
	# critical path).  Same kernels on the same inputs as the unpipelined order: results are bitwise unchanged.
	ORDERED_REF_MIN_Q = 8192	# from this batch size the pull-back reference walks its samples in cell order

	def set_samplers(self, data_fn, boundary_fn=None):
		"""data_fn(clock=None) -> (Q,3) samples; boundary_fn(clock=None, gs_dev=None) -> ((Qb,3) points, (Qb,3) normals), or (that pair,
		engine.Bins) when the sampler already ordered the batch (sample_box_surface_binned, on the sample grid gs_dev); clock: the device
		counter to draw with (None: the optimiser state's running clock); both must write persistent tensors"""
		self._samplers = (data_fn, boundary_fn)
		self._prep = None

	def _side_streams(self, census):
		if self._streams is None:
			self._streams = (torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream())
		main = torch.cuda.current_stream()
		return (main,) + (self._streams if census is None else (main, main, main))	# the census pass shares one counter: keep it serial

	def _prep_boundary(self, prep, clock=None, gs_dev=None):
		"""the boundary batch of an iteration, drawn and ordered on the CURRENT stream; gs_dev: the sample grid scale to bin with"""
		e = self.gv._engine
		boundary_fn = self._samplers[1]
		prep['gs_b'] = gs_dev
		res = boundary_fn(clock=clock, gs_dev=gs_dev)
		if isinstance(res[1], engine.Bins):	# the sampler drew and ordered the batch in one launch
			prep['boundary'], prep['bins_b'] = res
		else:
			prep['boundary'] = res
			prep['bins_b'] = e.bin_samples(res[0], True, tag='pb', gs_dev=gs_dev)

	def _prep_training(self, prep, s_ref, clock=None, gs_dev=None):
		"""the training batch on the CURRENT stream, and its RK4 pull-back reference on s_ref"""
		e = self.gv._engine
		data_fn = self._samplers[0]
		prep['gs_t'] = gs_dev
		data = prep['data'] = data_fn(clock) if clock is not None else data_fn()
		cur_stream = torch.cuda.current_stream()
		made = torch.cuda.Event()
		made.record(cur_stream)
		prep['bins'] = e.bin_samples(data, True, tag='pt', gs_dev=gs_dev)
		ev = torch.cuda.Event()
		ev.record(cur_stream)
		prep['ev'].append(ev)
		# the RK4 pull-back reference reads only the samples and the PREVIOUS field, which does not change during the phase: it
		# starts as soon as the samples exist and has until the adjoint kernel of the next iteration to finish — beside the hash
		# builds and the forward passes, off the critical path (it was 14 us of the longer of the two chains)
		# small batches are walked in their natural order (no need to wait for the hash); large ones in cell order, which keeps
		# the warps of the pull-back on neighbouring Gaussians (measured at N = Q = 64000: 419 -> 341 ms per step)
		Q = data.shape[0]
		ordered = Q >= self.ORDERED_REF_MIN_Q
		s_ref.wait_event(ev if ordered else made)
		with torch.cuda.stream(s_ref):
			if not ordered and (self._ident is None or self._ident.perm.shape[0] != Q):
				self._ident = engine.Bins(torch.arange(Q, dtype=torch.int32, device=data.device))
			ref_vor, ref_hel = self._tmp('ref_vor', (Q, 3)), self._tmp('ref_hel', (Q,))
			self.ref.velocity_field._engine.advected_vorticity(data, -self.ref.time_step, ref_vor, ref_hel, perm=prep['bins'] if ordered else self._ident)
			prep['ev_ref'] = torch.cuda.Event()
			prep['ev_ref'].record(s_ref)

	def _prepare(self, census=None, gs_dev=None):
		"""fork: samples + sample hash of the next iteration on the two side streams (not joined: see _join_prepared)"""
		main, s_fwd, s_bnd, s_ref = self._side_streams(census)
		fork = torch.cuda.Event()
		fork.record(main)
		prep = {'boundary': None, 'ev': [], 'gs_b': None, 'gs_t': None}
		if self._samplers[1] is not None:
			s_bnd.wait_event(fork)
			with torch.cuda.stream(s_bnd):
				self._prep_boundary(prep, gs_dev=gs_dev)
				ev = torch.cuda.Event()
				ev.record(s_bnd)
				prep['ev'].append(ev)
		s_fwd.wait_event(fork)
		with torch.cuda.stream(s_fwd):
			self._prep_training(prep, s_ref, gs_dev=gs_dev)
		self._prep = prep

	def _join_prepared(self, join_ref=True):
		"""join the prepared samples and their hashes; the pull-back reference only when asked (it is needed at the adjoint kernel
		of the next iteration, which waits for it itself — but a captured graph must end with every stream joined)"""
		main = torch.cuda.current_stream()
		for ev in self._prep['ev']:
			main.wait_event(ev)
		self._prep['ev'] = []
		if join_ref and self._prep.get('ev_ref') is not None:
			main.wait_event(self._prep['ev_ref'])
			self._prep['ev_ref'] = None

	def prime(self, census=None):
		"""prologue of a pipelined phase: prepare the samples of its first iteration (binned on the sample grid of iteration 0 when
		the sample grids run ahead of the hash)"""
		self._prepare(census, gs_dev=self.sample_grid())
		self._join_prepared()

	def iterate(self, data=None, boundary=None, census=None, join_all=True):
		"""
		One optimiser iteration.  Three independent chains run on three streams (fork / join by events, so a captured graph
		keeps the concurrency): the RK4 pull-back reference (previous field), the forward pass of the current field, and the
		whole boundary pass; they meet at the adjoint kernel and at the all-reduce.
		data=None: pipelined — use the samples prepared by prime() / the previous iteration and prepare the next ones;
		join_all=False leaves the next pull-back reference running into the next call (not for the last call of a captured graph).

		Pipelined with self.ahead (the default): the batches of iteration k + 1 are drawn and ordered WHILE iteration k runs — the
		boundary batch on its stream right behind the boundary gather, the training batch behind the training gather — instead of
		behind step k.  They need nothing from step k: the sample counter is the running clock + 1, and the samples are ordered on the
		sample grid of iteration k + 1 (sample_grid(): left by step k - 1, checked by step k against the hash it builds).  What
		remains on the cycle is step (+ hash, same launch) -> boundary forward -> boundary adjoint + gather.
		"""
		gv, e = self.gv, self.gv._engine
		cur = self.ref.velocity_field
		pipelined = data is None
		it = self._it
		self._it += 1
		parity = it & 1
		sgs_now, sgs_next = self.sample_grid(it), self.sample_grid(it + 1)
		prep = None
		if pipelined:
			prep = self._prep
			if prep is None or prep['ev']:
				raise _lib.GsrError('pipelined iterate() needs set_samplers() + prime() first')
			data, boundary = prep['data'], prep['boundary']
		ahead = pipelined and self.ahead
		Q, Qg = data.shape[0], data.shape[0] * self.world
		acc_w, lp_w, lpb_w = self.views[parity if self.peer else 0]
		acc_r, lp_r, lpb_r = self.reduced
		main, s_fwd, s_bnd, s_ref = self._side_streams(census)
		fork = torch.cuda.Event()
		fork.record(main)
		nxt = {'boundary': None, 'ev': [], 'gs_b': None, 'gs_t': None}
		ev_clock = None
		if ahead:	# the sample counter of the next batches, read before this iteration's step advances the clock
			s_ref.wait_event(fork)
			with torch.cuda.stream(s_ref):
				torch.add(self.stepper.clock, 1., out=self._clock1)
				ev_clock = torch.cuda.Event()
				ev_clock.record(s_ref)
		mask_b = 0
		srcs_b = []
		if boundary is not None:
			s_bnd.wait_event(fork)
			with torch.cuda.stream(s_bnd):
				bdata, bnormal = boundary
				Qb, Qbg = bdata.shape[0], bdata.shape[0] * self.world
				bins_b = prep['bins_b'] if pipelined else e.bin_samples(bdata, True, tag='b', gs_dev=sgs_now)
				perm_b, scs_b = bins_b
				valb = self._tmp('valb', (Qb, 3))
				e.forward(bdata, valb, None, accumulate=False, perm=bins_b)
				_, mask_b = e.backward_gather(bdata, perm_b, scs_b, valb, None, (0., self.boundary_lambda, 0., 0., 0., 0.),
											  {'normals': bnormal}, None, Q_norm=Qbg, tag='acc_b', acc=acc_w, loss_partials=lpb_w,
											  sample_gs=prep['gs_b'] if pipelined else sgs_now)
				srcs_b = [(lpb_r, self.nblkb, [0., 0., 0., self.boundary_lambda / Qbg, 0., 0., 0., 0.])]
				if census is not None:
					e.count_pairs(bdata, census.c, 2, True)
				done_b = torch.cuda.Event()
				done_b.record(s_bnd)
				if ahead:	# same stream, behind the consumers of the buffers it overwrites
					s_bnd.wait_event(ev_clock)
					self._prep_boundary(nxt, clock=self._clock1, gs_dev=sgs_next)
					ev = torch.cuda.Event()
					ev.record(s_bnd)
					nxt['ev'].append(ev)
		bins = prep['bins'] if pipelined else e.bin_samples(data, True, gs_dev=sgs_now)
		perm, scs = bins
		ref_vor, ref_hel = self._tmp('ref_vor', (Q, 3)), self._tmp('ref_hel', (Q,))
		val, grad = self._tmp('val', (Q, 3)), self._tmp('grad', (Q, 3, 3))
		binned = torch.cuda.Event()
		binned.record(main)
		s_fwd.wait_event(binned)
		with torch.cuda.stream(s_fwd):
			e.forward(data, val, grad, accumulate=False, perm=bins)
			done_f = torch.cuda.Event()
			done_f.record(s_fwd)
		if not pipelined:
			cur._engine.advected_vorticity(data, -self.ref.time_step, ref_vor, ref_hel, perm=bins)
		elif prep.get('ev_ref') is not None:	# the pull-back of these samples was launched when they were generated
			main.wait_event(prep['ev_ref'])
			prep['ev_ref'] = None
		main.wait_event(done_f)
		_, mask = e.backward_gather(data, perm, scs, val, grad, (0., 0., 0., self.w['vor'], self.w['hel'], self.w['div']),
									{'ref_vor': ref_vor, 'ref_hel': ref_hel}, None, Q_norm=Qg, acc=acc_w, loss_partials=lp_w,
									sample_gs=prep['gs_t'] if pipelined else sgs_now)
		srcs = [(lp_r, self.nblk, [self.w['vor'] / Qg, 0., self.w['div'] / Qg, 0., 0., 0., 0., 0.])]
		if census is not None:
			cur._engine.count_pairs(data, census.c, 5, True)
			e.count_pairs(data, census.c, 2, True)
		if ahead:	# the training batch of the next iteration: its buffers are free once this gather has read them
			gathered = torch.cuda.Event()
			gathered.record(main)
			s_fwd.wait_event(gathered)
			s_fwd.wait_event(ev_clock)
			with torch.cuda.stream(s_fwd):
				self._prep_training(nxt, s_ref, clock=self._clock1, gs_dev=sgs_next)
			main.wait_event(ev_clock)	# (long done) the step below advances the clock the counter was read from
		if boundary is not None:
			main.wait_event(done_b)
			mask |= mask_b
			srcs += srcs_b
		if self.peer:
			self.peer.sum(parity, self.stepper.state[:1])	# one kernel: handshake + sum of all ranks' buffers over NVLink
		elif self.world > 1:
			torch.distributed.all_reduce(self.flat)
		params = [p.detach() for p in gv._params()]
		if not pipelined or ahead:
			self.stepper.step(params, acc_r, mask, loss_srcs=srcs, rebuild=True)	# update + hash + packed records (one launch for N <= 1024)
			if ahead:
				self._prep = nxt
				self._join_prepared(join_ref=join_all)
			return
		self.stepper.step(params, acc_r, mask, loss_srcs=srcs)	# update; leaves iteration counter and grid_scale of the next iteration
		self._prepare(census, gs_dev=sgs_next)	# side streams: next samples + their hash ...
		self._rebuild()	# ... beside the Gaussian hash + packed records on this one
		self._join_prepared(join_ref=join_all)

	# ---- a phase: begin(), run_iterations() as often as needed, finish() -----------------------------------------------------------
	def begin(self, census=None):
		"""prepare the samples of the phase's first iteration"""
		self.prime(census)

	def run_iterations(self, n, census=None, use_graph=True):
		"""
		n pipelined iterations.  Iterations run in units of `unit` — one captured CUDA graph per projector (a replay costs ~5 us of
		launch overhead, tools/graph_probe.py, so several iterations share one; the peer-memory exchange alternates between two
		buffers, so a unit then holds whole pairs) — and a remainder eagerly.  The first unit of a new projector runs eagerly (it
		sizes every scratch buffer), then the graph is captured.
		"""
		unit = int(os.environ.get('GSR_GRAPH_UNIT', '0'))
		if unit <= 0:
			# whole pairs per captured unit: the exchange buffers and the sample grids alternate with the iteration's parity
			unit = self.unit or next((u for u in (10, 4, 2) if n % u == 0), 1)
		if self.peer and (unit % 2 or n % 2):
			raise ValueError('with the peer-memory exchange iterations run in pairs')
		if unit % 2:
			use_graph = False	# an odd number of iterations: no static parity to capture, run them one by one
		e = self.gv._engine

		def body(k_iters, cen=None):
			for k in range(k_iters):
				self.iterate(None, None, cen, join_all=(k == k_iters - 1))

		done = 0
		graphed = use_graph and census is None
		if graphed and self.graph is None and n >= 2 * unit:
			side = torch.cuda.Stream()
			side.wait_stream(torch.cuda.current_stream())
			with torch.cuda.stream(side):
				l0 = e.lib.gsr_launch_count()
				body(unit)	# eager warm-up (it counts)
				self.per_iter = e.lib.gsr_launch_count() - l0
				done += unit
			torch.cuda.current_stream().wait_stream(side)
			self.graph, self.unit = torch.cuda.CUDAGraph(), unit
			from .graphloop import capture_guard
			with capture_guard(), torch.cuda.graph(self.graph):
				body(unit)
			self.graph_launches -= self.per_iter	# the capture pass bumped the host counter without running anything
		while done < n:
			if graphed and self.graph is not None and n - done >= self.unit:
				self.graph.replay()
				self.graph_launches += self.per_iter	# kernels of this library inside one replay
				done += self.unit
			else:
				k = min(unit, n - done)
				body(k, census)
				done += k

	def evaluate_global(self, data, total=None, probe=None, census=None, fixed=False):
		"""the test losses over the WHOLE lattice: `data` is this rank's share, `total` the number of points of all ranks; fixed: see
		FusedProjector.evaluate (the pull-back reference of a fixed lattice is evaluated by the first test pass of the phase only)"""
		Q = data.shape[0]
		if self.peer:
			self.peer.check()	# (this call synchronises anyway) a lost peer is reported within check_iter iterations, not at the end of the phase
		self.check_sample_grid()
		sums = self.evaluate(data, probe=probe, fixed=fixed) * Q
		if census is not None:
			census.count(self.ref.velocity_field._engine, data, 5, lattice=True, executed=not self.reference_reused)
			census.count(self.gv._engine, data, 1, lattice=True)
		if self.lattice_world > 1:
			torch.distributed.all_reduce(sums)
		return sums / float(total or Q)

class LeapfrogTimestep:
	def __init__(self, n=10, dt=.02, iters=600, Qb=8192, test_res=128, boundary_lambda=10., rank=0, world=1, seed=42, check_iter=100, use_graph=True,
				 scaling='weak'):
		self.n, self.dt, self.iters, self.Qb, self.test_res, self.boundary_lambda = n, dt, iters, Qb, test_res, boundary_lambda
		self.rank, self.world, self.check_iter, self.use_graph, self.scaling = rank, world, check_iter, use_graph, scaling
		P, S, R, V, mgs, _ = synthetic_field(n, seed)
		self.params0 = (P, S, R, V)
		self.cur = make_fast3d(P, S, R, V, 5e-3, mgs)
		self.new = make_fast3d(P, S, R, V, 5e-3, mgs)
		self.N = self.cur.N
		dev = gsr3d.device
		self.lattice, total = shard_lattice(test_res, rank, world, dev, scaling)
		box = (0., 1.) * 3
		self.data_gen = advance3d.BoxSampler(*box)
		self.test_gen = advance3d.LatticeGenerator(*box, test_res, test_res, test_res, points=self.lattice, total=total)
		self.boundary_gen = advance3d.BoxSurfaceSampler(*box) if boundary_lambda else None
		# strong scaling: the training iterations run replicated (no exchange); only the lattice is shared
		self.train_world = world if scaling == 'weak' else 1
		self.last_test = None
		self.probe = None
		self._seeds = {id(self.cur): seed, id(self.new): seed + 1}	# each field keeps its own sample stream when it is the optimised one
		self._fields0 = (self.cur, self.new)

	@property
	def graph_launches(self):
		"""kernels of this library launched from inside graph replays so far (the host-side launch counter does not see them)"""
		return sum(fp.graph_launches for f in (self.cur, self.new) for fp in f.__dict__.get('_pipelines', {}).values())

	def reset(self, params=None, reset_clock=True):
		"""restore both fields to the given (default: initial) parameters and the initial roles (which object is the current field, which
		the one being optimised; with reset_clock the sample clocks too, so that the next step repeats the first) — used between
		timed steps and by the e2e path"""
		self.cur, self.new = self._fields0
		P, S, R, V = params if params is not None else [torch.as_tensor(a) for a in self.params0]
		with torch.no_grad():
			for f in (self.cur, self.new):
				for t, src in zip((f.positions, f.scalings, f.rotations, f.values), (P, S, R, V)):
					t.copy_(torch.as_tensor(src), non_blocking=True)
				f.zero_grad()
				if reset_clock:
					for fp in f.__dict__.get('_pipelines', {}).values():
						fp._drop_clock = True

	def step(self, census=None):
		"""one frame through the public API: advance3d.advance_frame (clone -> advect -> project -> swap -> the two output fields)"""
		cur, new = self.cur, self.new
		if census is not None:
			cur._engine.count_pairs(cur.positions.detach(), census.c, 4, True)	# the advect pass (RK4 positions only)
		hist = {}
		self.cur, self.new, _, (vor, div) = advance3d.advance_frame(
			cur, new, 0., 1., 0., 1., 0., 1., self.dt, self.data_gen, self.test_gen, boundary_generator=self.boundary_gen,
			boundary_lambda=self.boundary_lambda, max_epoch=self.iters, patience=10 ** 9, verbose=0, batch_size=self.Qb, check_iter=self.check_iter,
			rank=self.rank if self.train_world > 1 else 0, world=self.train_world, census=census, probe=self.probe, use_graph=self.use_graph,
			history=hist, lattice_world=self.world, sample_seed=self._seeds[id(new)])
		if hist.get('test'):
			t = hist['test'][-1]
			self.last_test = torch.tensor([t['loss_vor'], t['loss_hel'], t['loss_div']])
		return vor, div
